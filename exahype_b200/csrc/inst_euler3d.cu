// Committed instantiations: compressible Euler, 3-D (5 unknowns), fp64 and fp32.
// BASELINE.json config C3/C5: 8x8x8 patches + 1 halo.  Default kernel: warp per patch (fv3d_pair_kernel.cuh), 8 warps per
// CTA, 4000-byte planes streamed through a 4-deep TMA ring per warp.  -DEXAHYPE_3D_PAIR=0 builds the
// previous default instead, plane marching by groups of three warps (fv3d_march_kernel.cuh: 5 groups per CTA, 5-deep
// ring), which still serves 4x4x4 patches.  Alternative (EXAHYPE_FLAG_KERNEL_CELL): the thread-per-cell kernel, one
// patch per tile, 512 threads, 40 000-byte tiles by TMA.
#include <vector>

#include "fv_registry.h"

namespace exahype {
namespace {
using E3 = EulerPhysics<3, 5, 0>;
constexpr int EU = EXAHYPE_MODEL_EULER, F64 = EXAHYPE_DTYPE_F64, F32 = EXAHYPE_DTYPE_F32;

#ifndef EXAHYPE_3D_NG
#define EXAHYPE_3D_NG 5   // warp groups per CTA of the plane-marching kernel for 8^3 patches
#endif
#ifndef EXAHYPE_3D_R
#define EXAHYPE_3D_R 5    // planes in each group's TMA ring
#endif

#ifndef EXAHYPE_3D_PAIR
#define EXAHYPE_3D_PAIR 1   // 8^3 patches: warp-per-patch kernel (fv3d_pair_kernel.cuh) instead of the three-warp groups
#endif
#ifndef EXAHYPE_3D_NW
#define EXAHYPE_3D_NW 8     // warps (= patches in flight) per CTA of the warp-per-patch kernel
#endif
#ifndef EXAHYPE_3D_NW_F32
#define EXAHYPE_3D_NW_F32 8   // fp32: two CTAs per SM fit (16 warps), measured faster than one CTA of 12 warps
#endif
#ifndef EXAHYPE_3D_PR
#define EXAHYPE_3D_PR 4     // planes in each warp's TMA ring
#endif
#if EXAHYPE_3D_PAIR
#ifndef EXAHYPE_3D_SB
#define EXAHYPE_3D_SB 1     // output staging buffers per warp (one: leaves room for the fourth ring slot)
#endif
using Main8d = Pair3dFamily<E3, double, 8, 1, EXAHYPE_3D_NW, EXAHYPE_3D_PR, EXAHYPE_3D_SB>;
using Main8f = Pair3dFamily<E3, float, 8, 1, EXAHYPE_3D_NW_F32, EXAHYPE_3D_PR, EXAHYPE_3D_SB>;
#else
using Main8d = March3dFamily<E3, double, 8, 1, EXAHYPE_3D_NG, EXAHYPE_3D_R, 1>;
using Main8f = March3dFamily<E3, float, 8, 1, 5, 5, 1>;
#endif

const std::vector<FvEntry>& entries() {
  static const std::vector<FvEntry> v = {
      //          plane marching: phys, T, P, H, NG groups, R planes, CTAs/SM | thread per cell: phys, T, dim, P, H, G, NT, CTAs/SM
      march_entry<Main8d, CellFamily<E3, double, 3, 8, 1, 1, 512, 1>>(EU, F64, 3, 8, 1, 5, 0),
      march_entry<Main8f, CellFamily<E3, float, 3, 8, 1, 1, 512, 1>>(EU, F32, 3, 8, 1, 5, 0),
      march_entry<March3dFamily<E3, double, 4, 1, 6, 6, 1>, CellFamily<E3, double, 3, 4, 1, 4, 256, 2>>(EU, F64, 3, 4, 1, 5, 0),
      march_entry<March3dFamily<E3, float, 4, 1, 6, 6, 1>, CellFamily<E3, float, 3, 4, 1, 4, 256, 2>>(EU, F32, 3, 4, 1, 5, 0),
  };
  return v;
}
}  // namespace

FvEntryList euler3d_entries() { return {entries().data(), (int)entries().size()}; }
}  // namespace exahype
