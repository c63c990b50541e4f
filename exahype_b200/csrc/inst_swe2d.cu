// Committed instantiations: 2-D shallow water (h, hu, hv | bathymetry), fp64 and fp32.
// BASELINE.json config C4: 32x32 patches + 1 halo -- row marching, one patch per warp (alternative: thread-per-cell,
// 1024 interior cells, two per thread).
#include "fv_registry.h"

namespace exahype {
namespace {
using SW = SwePhysics<3, 1>;

const FvEntry kEntries[] = {
    // row-marching kernel (default): WPC warps per CTA, MINB, PF rows of register prefetch | thread-per-cell kernel: G, NT, MINB
    EXAHYPE_FV2D_ENTRY(EXAHYPE_MODEL_SWE, EXAHYPE_DTYPE_F64, SW, double, 32, 1, 4, 4, 3, 1, 512, 1),
    EXAHYPE_FV2D_ENTRY(EXAHYPE_MODEL_SWE, EXAHYPE_DTYPE_F32, SW, float, 32, 1, 4, 4, 3, 1, 512, 1),
    EXAHYPE_FV2D_ENTRY(EXAHYPE_MODEL_SWE, EXAHYPE_DTYPE_F64, SW, double, 16, 1, 4, 4, 2, 1, 256, 2),
    EXAHYPE_FV2D_ENTRY(EXAHYPE_MODEL_SWE, EXAHYPE_DTYPE_F32, SW, float, 16, 1, 4, 4, 3, 1, 256, 2),
};
}  // namespace

FvEntryList swe2d_entries() { return {kEntries, (int)(sizeof(kEntries) / sizeof(kEntries[0]))}; }
}  // namespace exahype
