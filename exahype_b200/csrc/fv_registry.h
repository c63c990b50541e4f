// Table of committed kernel instantiations; each translation unit inst_*.cu contributes one array.
#pragma once

#include <cuda_runtime.h>

#include <type_traits>

#include "../../include/exahype_cuda.h"
#include "fv_patch_kernel.cuh"
#include "fv3d_march_kernel.cuh"
#include "fv3d_pair_kernel.cuh"
#include "fv2d_march_kernel.cuh"

namespace exahype {

using FvLaunchFn = cudaError_t (*)(const void*, void*, long long, double, void*, void*, cudaStream_t, const FvGatherRaw*);
using FvPrepareFn = cudaError_t (*)(FvLaunchInfo*, long long);

struct FvEntry {
  exahype_fv_config cfg;      // flags == 0; the variant is picked from the caller's flags
  FvLaunchFn launch[4];       // index = (DISSIPATION_ALL ? 1 : 0) | (OUTPUT_UNHALOED ? 2 : 0)
  FvPrepareFn prepare[4];
  // second kernel for the same shape (the thread-per-cell kernel where `launch` is the plane-marching one);
  // selected with EXAHYPE_FLAG_KERNEL_CELL, null when there is only one kernel
  FvLaunchFn alt_launch[4];
  FvPrepareFn alt_prepare[4];
  // the default kernel instantiated for the CellData form (per-patch pointers / time steps): exahype_cuda_fv_step_cell_data
  FvLaunchFn gather_launch[4];
  // un-haloed output without the auxiliary variables (EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY), index = DISSIPATION_ALL;
  // null when the shape's default kernel has no such form (shapes without auxiliary variables never need one)
  FvLaunchFn unknowns_launch[2];
  FvPrepareFn unknowns_prepare[2];
};

struct FvEntryList {
  const FvEntry* entries;
  int count;
};

FvEntryList euler2d_entries();
FvEntryList euler3d_entries();
FvEntryList swe2d_entries();
FvEntryList fast_entries();     // inst_fast.cu: ArithFast + -fmad=true, looked up under EXAHYPE_FLAG_FAST_ARITHMETIC

// A committed shape is described by up to three launcher families, each a template over (DISSIPATION_ALL, UNHALOED):
// Main (default kernel, dense batch), Gather (same kernel, CellData form) and Alt (second kernel, or NoKernel).
template <bool DA, bool UH>
struct NoKernel {
  static constexpr bool exists = false;
  static cudaError_t prepare(FvLaunchInfo*, long long) { return cudaErrorNotSupported; }
  static cudaError_t launch(const void*, void*, long long, double, void*, void*, cudaStream_t, const FvGatherRaw*) {
    return cudaErrorNotSupported;
  }
};
template <class L, class = void> struct launcher_exists { static constexpr bool value = true; };
template <class L> struct launcher_exists<L, std::enable_if_t<!L::exists>> { static constexpr bool value = false; };

template <template <bool, bool> class Main, template <bool, bool> class Gather, template <bool, bool> class Alt>
inline FvEntry make_entry(int model, int dtype, int dim, int P, int H, int nr, int na) {
  constexpr bool alt = launcher_exists<Alt<false, false>>::value;
  FvEntry e = {};
  e.cfg = {model, dtype, dim, P, H, nr, na, 0u};
  FvLaunchFn ml[4] = {&Main<false, false>::launch, &Main<true, false>::launch, &Main<false, true>::launch, &Main<true, true>::launch};
  FvPrepareFn mp[4] = {&Main<false, false>::prepare, &Main<true, false>::prepare, &Main<false, true>::prepare, &Main<true, true>::prepare};
  FvLaunchFn gl[4] = {&Gather<false, false>::launch, &Gather<true, false>::launch, &Gather<false, true>::launch, &Gather<true, true>::launch};
  FvLaunchFn al[4] = {&Alt<false, false>::launch, &Alt<true, false>::launch, &Alt<false, true>::launch, &Alt<true, true>::launch};
  FvPrepareFn ap[4] = {&Alt<false, false>::prepare, &Alt<true, false>::prepare, &Alt<false, true>::prepare, &Alt<true, true>::prepare};
  for (int i = 0; i < 4; ++i) {
    e.launch[i] = ml[i];
    e.prepare[i] = mp[i];
    e.gather_launch[i] = gl[i];
    e.alt_launch[i] = alt ? al[i] : nullptr;
    e.alt_prepare[i] = alt ? ap[i] : nullptr;
  }
  return e;
}

// ---- launcher families for the three kernel templates (RusanovUpdate is the update functor of every committed shape)
// thread per cell: G patches per tile, NT threads, MINB CTAs per SM
template <class Phys, typename T, int DIM, int P, int H, int G, int NT, int MINB>
struct CellFamily {
  template <bool DA, bool UH> using Dense = FvLauncher<FvKernelConfig<Phys, RusanovUpdate, T, DIM, P, H, G, NT, MINB, DA, UH, false>>;
  template <bool DA, bool UH> using Gather = FvLauncher<FvKernelConfig<Phys, RusanovUpdate, T, DIM, P, H, G, NT, MINB, DA, UH, true>>;
};
// 3-D plane marching: NG groups per CTA, ring of R planes, MINB CTAs per SM
template <class Phys, typename T, int P, int H, int NG, int R, int MINB>
struct March3dFamily {
  template <bool DA, bool UH> using Dense = Fv3dMarchLauncher<Fv3dMarchConfig<Phys, RusanovUpdate, T, P, H, NG, R, MINB, DA, UH, false>>;
  template <bool DA, bool UH> using Gather = Fv3dMarchLauncher<Fv3dMarchConfig<Phys, RusanovUpdate, T, P, H, NG, R, MINB, DA, UH, true>>;
  static constexpr bool HAS_UNKNOWNS = false;
};
// 3-D warp-per-patch marching (8x8x8 patches): NW warps per CTA, ring of R planes per warp
template <class Phys, typename T, int P, int H, int NW, int R, int SB = 2>
struct Pair3dFamily {
  template <bool DA, bool UH> using Dense = Fv3dPairLauncher<Fv3dPairConfig<Phys, RusanovUpdate, T, P, H, NW, R, DA, UH, false, SB>>;
  template <bool DA, bool UH> using Gather = Fv3dPairLauncher<Fv3dPairConfig<Phys, RusanovUpdate, T, P, H, NW, R, DA, UH, true, SB>>;
  static constexpr bool HAS_UNKNOWNS = false;
};
// 2-D row marching: WPC warps per CTA, MINB CTAs per SM, PF rows of register prefetch
template <class Phys, typename T, int P, int H, int WPC, int MINB, int PF>
struct March2dFamily {
  static constexpr int WIDE = Fv2dVec<T, Phys::NR + Phys::NA>::WIDE, NARROW = Fv2dVec<T, Phys::NR + Phys::NA>::NARROW;
  template <bool DA, bool UH>
  using Dense = Fv2dMarchDispatch<Fv2dMarchConfig<Phys, RusanovUpdate, T, P, H, WPC, MINB, DA, UH, WIDE, PF, false>,
                                  Fv2dMarchConfig<Phys, RusanovUpdate, T, P, H, WPC, MINB, DA, UH, NARROW, PF, false>>;
  template <bool DA, bool UH>
  using Gather = Fv2dMarchLauncher<Fv2dMarchConfig<Phys, RusanovUpdate, T, P, H, WPC, MINB, DA, UH, NARROW, PF, true>>;
  template <bool DA>
  using Unknowns = Fv2dMarchDispatch<Fv2dMarchConfig<Phys, RusanovUpdate, T, P, H, WPC, MINB, DA, true, WIDE, PF, false, true>,
                                     Fv2dMarchConfig<Phys, RusanovUpdate, T, P, H, WPC, MINB, DA, true, NARROW, PF, false, true>>;
  static constexpr bool HAS_UNKNOWNS = Phys::NA > 0;
};
template <class March>
inline void add_unknowns_form(FvEntry& e) {
  if constexpr (March::HAS_UNKNOWNS) {
    e.unknowns_launch[0] = &March::template Unknowns<false>::launch;
    e.unknowns_launch[1] = &March::template Unknowns<true>::launch;
    e.unknowns_prepare[0] = &March::template Unknowns<false>::prepare;
    e.unknowns_prepare[1] = &March::template Unknowns<true>::prepare;
  }
}

// one kernel for the shape (thread per cell)
template <class Cell>
inline FvEntry cell_entry(int model, int dtype, int dim, int P, int H, int nr, int na) {
  return make_entry<Cell::template Dense, Cell::template Gather, NoKernel>(model, dtype, dim, P, H, nr, na);
}
// marching kernel by default, thread-per-cell kernel behind EXAHYPE_FLAG_KERNEL_CELL
template <class March, class Cell>
inline FvEntry march_entry(int model, int dtype, int dim, int P, int H, int nr, int na) {
  FvEntry e = make_entry<March::template Dense, March::template Gather, Cell::template Dense>(model, dtype, dim, P, H, nr, na);
  add_unknowns_form<March>(e);
  return e;
}

// marching kernel only (no thread-per-cell alternative: the shape's tile does not fit shared memory, or none is wanted)
template <class March>
inline FvEntry march_only_entry(int model, int dtype, int dim, int P, int H, int nr, int na) {
  FvEntry e = make_entry<March::template Dense, March::template Gather, NoKernel>(model, dtype, dim, P, H, nr, na);
  add_unknowns_form<March>(e);
  return e;
}

}  // namespace exahype
