// Committed instantiations with the opt-in fast arithmetic (EXAHYPE_FLAG_FAST_ARITHMETIC): the headline shapes of
// BASELINE.json with the ArithFast policy of physics.cuh (branch-free reciprocal / square root), and THIS translation unit
// is compiled with -fmad=true (exahype_b200/build.py), so multiply-add pairs contract.  Same statement order, same
// kernels; results within 1e-12 relative of the reference arithmetic (not bitwise).  14 % fewer instructions in the 3-D
// kernel: with them the single output staging buffer per warp becomes the limit (the next plane's update waits for the
// previous plane's bulk store to have read it), so the 8^3 kernel runs a 3-deep TMA ring with TWO staging buffers here.
//
// Every physics type in this file carries ArithFast, so no kernel instantiated here has the name of one instantiated in
// the contraction-free units.
#include <vector>

#include "fv_registry.h"

namespace exahype {
namespace {
using E3 = EulerPhysics<3, 5, 0, ArithFast>;
using E2 = EulerPhysics<2, 4, 0, ArithFast>;
using SW = SwePhysics<3, 1, ArithFast>;
constexpr int EU = EXAHYPE_MODEL_EULER, SWE = EXAHYPE_MODEL_SWE, F64 = EXAHYPE_DTYPE_F64, F32 = EXAHYPE_DTYPE_F32;

#ifndef EXAHYPE_FAST_3D_NW
#define EXAHYPE_FAST_3D_NW 8     // warps (= patches in flight) per CTA
#endif
#ifndef EXAHYPE_FAST_3D_PR
#define EXAHYPE_FAST_3D_PR 3     // planes in each warp's TMA ring
#endif
#ifndef EXAHYPE_FAST_3D_SB
#define EXAHYPE_FAST_3D_SB 2     // output staging buffers per warp
#endif

const std::vector<FvEntry>& entries() {
  static const std::vector<FvEntry> v = {
      march_only_entry<Pair3dFamily<E3, double, 8, 1, EXAHYPE_FAST_3D_NW, EXAHYPE_FAST_3D_PR, EXAHYPE_FAST_3D_SB>>(EU, F64, 3, 8, 1, 5, 0),   // C3 / C5
      march_only_entry<March2dFamily<E2, double, 16, 1, 1, 16, 2>>(EU, F64, 2, 16, 1, 4, 0),                                 // C2
      march_only_entry<March2dFamily<SW, double, 32, 1, 1, 16, 3>>(SWE, F64, 2, 32, 1, 3, 1),                                // C4
      march_only_entry<March2dFamily<SW, float, 32, 1, 1, 24, 4>>(SWE, F32, 2, 32, 1, 3, 1),                                 // C4 fp32 (contraction only)
  };
  return v;
}
}  // namespace

FvEntryList fast_entries() { return {entries().data(), (int)entries().size()}; }
}  // namespace exahype
