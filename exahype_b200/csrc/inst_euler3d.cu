// Committed instantiations: compressible Euler, 3-D (5 unknowns), fp64 and fp32.
// BASELINE.json config C3/C5: 8x8x8 patches + 1 halo -- one patch per tile, 512 threads = one per interior cell,
// 40 000-byte tiles brought in by TMA bulk copies.
#include "fv_registry.h"

namespace exahype {
namespace {
using E3 = EulerPhysics<3, 5, 0>;

const FvEntry kEntries[] = {
    //                model                dtype              phys T      D  P  H  G   NT  MINB
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E3, double, 3, 8, 1, 1, 512, 1),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F32, E3, float, 3, 8, 1, 1, 512, 1),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, E3, double, 3, 4, 1, 4, 256, 2),
    EXAHYPE_FV_ENTRY(EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F32, E3, float, 3, 4, 1, 4, 256, 2),
};
}  // namespace

FvEntryList euler3d_entries() { return {kEntries, (int)(sizeof(kEntries) / sizeof(kEntries[0]))}; }
}  // namespace exahype
