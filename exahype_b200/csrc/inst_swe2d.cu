// Committed instantiations: 2-D shallow water (h, hu, hv | bathymetry), fp64 and fp32.
// BASELINE.json config C4: 32x32 patches + 1 halo -- row marching, one patch per warp (alternative: thread-per-cell,
// 1024 interior cells, two per thread).
#include <vector>

#include "fv_registry.h"

namespace exahype {
namespace {
using SW = SwePhysics<3, 1>;
using SWS = SweSourcePhysics<3, 3>;      // + bathymetry source term, aux = (b, db/dx, db/dy)   
constexpr int SWE = EXAHYPE_MODEL_SWE, SWES = EXAHYPE_MODEL_SWE_SOURCE, F64 = EXAHYPE_DTYPE_F64, F32 = EXAHYPE_DTYPE_F32;

#ifndef EXAHYPE_SWE32D_WPC
#define EXAHYPE_SWE32D_WPC 1    // warps per CTA / CTAs per SM of the fp64 32x32 row-marching kernel (one warp per CTA: inst_euler2d.cu)
#endif
#ifndef EXAHYPE_SWE32D_MINB
#define EXAHYPE_SWE32D_MINB 16
#endif
#ifndef EXAHYPE_SWE32F_WPC
#define EXAHYPE_SWE32F_WPC 1
#endif
#ifndef EXAHYPE_SWE32F_MINB
#define EXAHYPE_SWE32F_MINB 24  // CTAs (of one warp) per SM of the fp32 32x32 row-marching kernel: 24 warps at <= 85 registers
#endif
#ifndef EXAHYPE_SWE32F_PF
#define EXAHYPE_SWE32F_PF 4     // its register prefetch distance (rows); measured 0.389 -> 0.370 ms on C4 fp32 with 6 CTAs + 4 rows
#endif

const std::vector<FvEntry>& entries() {
  static const std::vector<FvEntry> v = {
      //          row marching: phys, T, P, H, warps/CTA, CTAs/SM, PF | thread per cell: phys, T, dim, P, H, G, NT, CTAs/SM
      march_entry<March2dFamily<SW, double, 32, 1, EXAHYPE_SWE32D_WPC, EXAHYPE_SWE32D_MINB, 3>, CellFamily<SW, double, 2, 32, 1, 1, 512, 1>>(SWE, F64, 2, 32, 1, 3, 1),
      march_entry<March2dFamily<SW, float, 32, 1, EXAHYPE_SWE32F_WPC, EXAHYPE_SWE32F_MINB, EXAHYPE_SWE32F_PF>, CellFamily<SW, float, 2, 32, 1, 1, 512, 1>>(SWE, F32, 2, 32, 1, 3, 1),
      march_entry<March2dFamily<SW, double, 16, 1, 1, 16, 2>, CellFamily<SW, double, 2, 16, 1, 1, 256, 2>>(SWE, F64, 2, 16, 1, 3, 1),
      march_entry<March2dFamily<SW, float, 16, 1, 1, 16, 3>, CellFamily<SW, float, 2, 16, 1, 1, 256, 2>>(SWE, F32, 2, 16, 1, 3, 1),
      // with the source statement: 48-byte cells (6 fp64 values) go through 128-bit accesses
      march_only_entry<March2dFamily<SWS, double, 32, 1, 1, 12, 2>>(SWES, F64, 2, 32, 1, 3, 3),     // (a thread-per-cell tile would not fit)
      march_entry<March2dFamily<SWS, float, 32, 1, 1, 16, 3>, CellFamily<SWS, float, 2, 32, 1, 1, 512, 1>>(SWES, F32, 2, 32, 1, 3, 3),
      march_entry<March2dFamily<SWS, double, 16, 1, 1, 12, 2>, CellFamily<SWS, double, 2, 16, 1, 1, 256, 2>>(SWES, F64, 2, 16, 1, 3, 3),
      cell_entry<CellFamily<SWS, double, 2, 4, 1, 16, 256, 2>>(SWES, F64, 2, 4, 1, 3, 3),
  };
  return v;
}
}  // namespace

FvEntryList swe2d_entries() { return {entries().data(), (int)entries().size()}; }
}  // namespace exahype
