// Committed instantiations: compressible Euler, 2-D (4 unknowns), fp64 and fp32.
//   P = 3   BASELINE.json config C1 (3x3 + 1 halo, 1 000 patches): 28 patches per tile fill 252 threads
//   P = 16  config C2 (16x16 + 1 halo, 65 536 patches): one patch per tile, one thread per interior cell
//   P = 4 with 5 + 5 variables is the shape of the reference's committed kernel ("Unit test/test.cpp":4-8)
#include "fv_registry.h"

namespace exahype {
namespace {
using E2 = EulerPhysics<2, 4, 0>;
using E2ref = EulerPhysics<2, 5, 5>;

const FvEntry kEntries[] = {
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E2, double, 2, 16, 1, 1, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F32, E2, float, 2, 16, 1, 1, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E2, double, 2, 3, 1, 28, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F32, E2, float, 2, 3, 1, 28, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E2, double, 2, 4, 1, 16, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E2, double, 2, 8, 1, 4, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E2ref, double, 2, 4, 1, 8, 128, 2),
};
}  // namespace

FvEntryList euler2d_entries() { return {kEntries, (int)(sizeof(kEntries) / sizeof(kEntries[0]))}; }
}  // namespace exahype
