// Committed instantiations: compressible Euler, 3-D (5 unknowns), fp64 and fp32.
// BASELINE.json config C3/C5: 8x8x8 patches + 1 halo.  Default kernel: plane marching (fv3d_march_kernel.cuh), 4 groups of
// 3 warps per CTA, 4000-byte planes streamed through a 5-deep TMA ring.  Alternative (EXAHYPE_FLAG_KERNEL_CELL): the
// thread-per-cell kernel, one patch per tile, 512 threads, 40 000-byte tiles by TMA.
#include "fv_registry.h"

namespace exahype {
namespace {
using E3 = EulerPhysics<3, 5, 0>;

#ifndef EXAHYPE_3D_NG
#define EXAHYPE_3D_NG 5   // warp groups per CTA of the plane-marching kernel for 8^3 patches
#endif

const FvEntry kEntries[] = {
    // plane-marching kernel (default): NG groups per CTA, ring of R planes | thread-per-cell kernel: G, NT, MINB
    //                  model                dtype              phys T      P  H  NG R MINB | G   NT  MINB
    EXAHYPE_FV3D_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E3, double, 8, 1, EXAHYPE_3D_NG, 5, 1, 1, 512, 1),
    EXAHYPE_FV3D_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F32, E3, float, 8, 1, 5, 5, 1, 1, 512, 1),
    EXAHYPE_FV3D_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E3, double, 4, 1, 6, 6, 1, 4, 256, 2),
    EXAHYPE_FV3D_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F32, E3, float, 4, 1, 6, 6, 1, 4, 256, 2),
};
}  // namespace

FvEntryList euler3d_entries() { return {kEntries, (int)(sizeof(kEntries) / sizeof(kEntries[0]))}; }
}  // namespace exahype
