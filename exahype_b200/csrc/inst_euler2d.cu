// Committed instantiations: compressible Euler, 2-D (4 unknowns), fp64 and fp32.
//   P = 3   BASELINE.json config C1 (3x3 + 1 halo, 1 000 patches): thread per cell, 28 patches per tile fill 252 threads
//   P = 16  config C2 (16x16 + 1 halo, 65 536 patches): row marching (fv2d_march_kernel.cuh), two patches per warp;
//           alternative (EXAHYPE_FLAG_KERNEL_CELL): one patch per tile, one thread per interior cell
//   P = 4 with 5 + 5 variables is the shape of the reference's committed kernel ("Unit test/test.cpp":4-8)
#include <vector>

#include "fv_registry.h"

namespace exahype {
namespace {
using E2 = EulerPhysics<2, 4, 0>;
using E2ref = EulerPhysics<2, 5, 5>;
constexpr int EU = EXAHYPE_MODEL_EULER, F64 = EXAHYPE_DTYPE_F64, F32 = EXAHYPE_DTYPE_F32;

#ifndef EXAHYPE_2D_MINB16
#define EXAHYPE_2D_MINB16 16  // CTAs per SM of the same kernel
#endif
#ifndef EXAHYPE_2D_WPC16
// warps per CTA of the same kernel.  ONE: a CTA's slot (registers, shared memory) is only handed to the next CTA when
// its last warp has finished, so with four warps per CTA three of them idle at the end of every CTA, and four start-up
// waits (a full DRAM latency, 14 % of all warp samples) coincide.  Measured 4 -> 1 warps per CTA (x 4 CTAs per SM):
// C2 0.2199 -> 0.2114 ms, C4 fp32 0.3995 -> 0.3696 ms, sustained C4 0.802 -> 0.785 ms; 2 and 8 warps lie in between /
// behind (profiles/r02_2d_warps_per_cta.txt).
#define EXAHYPE_2D_WPC16 1
#endif
#ifndef EXAHYPE_2D_PF16
#define EXAHYPE_2D_PF16 2   // register prefetch distance (rows) of the fp64 16x16 row-marching kernel
#endif

const std::vector<FvEntry>& entries() {
  static const std::vector<FvEntry> v = {
      //          row marching: phys, T, P, H, warps/CTA (one, see above), CTAs/SM, PF | thread per cell: phys, T, dim, P, H, G, NT, CTAs/SM
      march_entry<March2dFamily<E2, double, 16, 1, EXAHYPE_2D_WPC16, EXAHYPE_2D_MINB16, EXAHYPE_2D_PF16>, CellFamily<E2, double, 2, 16, 1, 1, 256, 2>>(EU, F64, 2, 16, 1, 4, 0),
      march_entry<March2dFamily<E2, float, 16, 1, 1, 16, 3>, CellFamily<E2, float, 2, 16, 1, 1, 256, 2>>(EU, F32, 2, 16, 1, 4, 0),
      march_entry<March2dFamily<E2, double, 8, 1, 1, 16, 2>, CellFamily<E2, double, 2, 8, 1, 4, 256, 2>>(EU, F64, 2, 8, 1, 4, 0),
      cell_entry<CellFamily<E2, double, 2, 3, 1, 28, 256, 2>>(EU, F64, 2, 3, 1, 4, 0),
      cell_entry<CellFamily<E2, float, 2, 3, 1, 28, 256, 2>>(EU, F32, 2, 3, 1, 4, 0),
      cell_entry<CellFamily<E2, double, 2, 4, 1, 16, 256, 2>>(EU, F64, 2, 4, 1, 4, 0),
      cell_entry<CellFamily<E2ref, double, 2, 4, 1, 8, 128, 2>>(EU, F64, 2, 4, 1, 5, 5),
  };
  return v;
}
}  // namespace

FvEntryList euler2d_entries() { return {entries().data(), (int)entries().size()}; }
}  // namespace exahype
