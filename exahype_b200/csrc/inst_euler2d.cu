// Committed instantiations: compressible Euler, 2-D (4 unknowns), fp64 and fp32.
//   P = 3   BASELINE.json config C1 (3x3 + 1 halo, 1 000 patches): 28 patches per tile fill 252 threads
//   P = 16  config C2 (16x16 + 1 halo, 65 536 patches): row marching (fv2d_march_kernel.cuh), two patches per warp;
//           alternative (EXAHYPE_FLAG_KERNEL_CELL): one patch per tile, one thread per interior cell
//   P = 4 with 5 + 5 variables is the shape of the reference's committed kernel ("Unit test/test.cpp":4-8)
#include "fv_registry.h"

namespace exahype {
namespace {
using E2 = EulerPhysics<2, 4, 0>;
using E2ref = EulerPhysics<2, 5, 5>;

const FvEntry kEntries[] = {
    // row-marching kernel (default): WPC warps per CTA, MINB, PF rows of register prefetch | thread-per-cell kernel: G, NT, MINB
    EXAHYPE_FV2D_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E2, double, 16, 1, 4, 4, 2, 1, 256, 2),
    EXAHYPE_FV2D_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F32, E2, float, 16, 1, 4, 4, 3, 1, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E2, double, 2, 3, 1, 28, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F32, E2, float, 2, 3, 1, 28, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E2, double, 2, 4, 1, 16, 256, 2),
    EXAHYPE_FV2D_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E2, double, 8, 1, 4, 4, 2, 4, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E2ref, double, 2, 4, 1, 8, 128, 2),
};
}  // namespace

FvEntryList euler2d_entries() { return {kEntries, (int)(sizeof(kEntries) / sizeof(kEntries[0]))}; }
}  // namespace exahype
