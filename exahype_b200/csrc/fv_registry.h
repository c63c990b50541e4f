// Table of committed kernel instantiations; each translation unit inst_*.cu contributes one array.
#pragma once

#include <cuda_runtime.h>

#include "../../include/exahype_cuda.h"
#include "fv_patch_kernel.cuh"
#include "fv3d_march_kernel.cuh"
#include "fv2d_march_kernel.cuh"

namespace exahype {

using FvLaunchFn = cudaError_t (*)(const void*, void*, long long, double, void*, void*, cudaStream_t);
using FvPrepareFn = cudaError_t (*)(FvLaunchInfo*, long long);

struct FvEntry {
  exahype_fv_config cfg;      // flags == 0; the variant is picked from the caller's flags
  FvLaunchFn launch[4];       // index = (DISSIPATION_ALL ? 1 : 0) | (OUTPUT_UNHALOED ? 2 : 0)
  FvPrepareFn prepare[4];
  // second kernel for the same shape (the thread-per-cell kernel where `launch` is the plane-marching one);
  // selected with EXAHYPE_FLAG_KERNEL_CELL, null when there is only one kernel
  FvLaunchFn alt_launch[4];
  FvPrepareFn alt_prepare[4];
};

struct FvEntryList {
  const FvEntry* entries;
  int count;
};

FvEntryList euler2d_entries();
FvEntryList euler3d_entries();
FvEntryList swe2d_entries();

#define EXAHYPE_FV_CFG(PHYS, T, DIM, P, H, G, NT, MINB, DA, UH) \
  ::exahype::FvKernelConfig<PHYS, ::exahype::RusanovUpdate, T, DIM, P, H, G, NT, MINB, DA, UH>

// one committed shape = four kernels (dissipation var0|all  x  output haloed|un-haloed)
#define EXAHYPE_FV_ENTRY(MODEL, DTYPE, PHYS, T, DIM, P, H, G, NT, MINB)                                   \
  {                                                                                                       \
    {MODEL, DTYPE, DIM, P, H, PHYS::NR, PHYS::NA, 0u},                                                    \
        {&::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, DIM, P, H, G, NT, MINB, false, false)>::launch,   \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, DIM, P, H, G, NT, MINB, true, false)>::launch,    \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, DIM, P, H, G, NT, MINB, false, true)>::launch,    \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, DIM, P, H, G, NT, MINB, true, true)>::launch},    \
        {&::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, DIM, P, H, G, NT, MINB, false, false)>::prepare,  \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, DIM, P, H, G, NT, MINB, true, false)>::prepare,   \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, DIM, P, H, G, NT, MINB, false, true)>::prepare,   \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, DIM, P, H, G, NT, MINB, true, true)>::prepare},   \
        {nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}                        \
  }

#define EXAHYPE_MARCH_CFG(PHYS, T, P, H, NG, R, MINB, DA, UH) \
  ::exahype::Fv3dMarchConfig<PHYS, ::exahype::RusanovUpdate, T, P, H, NG, R, MINB, DA, UH>

// 3-D shape served by the plane-marching kernel, with the thread-per-cell kernel (G, NT, MINB_CELL) as alternative
#define EXAHYPE_FV3D_ENTRY(MODEL, DTYPE, PHYS, T, P, H, NG, R, MINB, G, NT, MINB_CELL)                           \
  {                                                                                                       \
    {MODEL, DTYPE, 3, P, H, PHYS::NR, PHYS::NA, 0u},                                                      \
        {&::exahype::Fv3dMarchLauncher<EXAHYPE_MARCH_CFG(PHYS, T, P, H, NG, R, MINB, false, false)>::launch,     \
         &::exahype::Fv3dMarchLauncher<EXAHYPE_MARCH_CFG(PHYS, T, P, H, NG, R, MINB, true, false)>::launch,      \
         &::exahype::Fv3dMarchLauncher<EXAHYPE_MARCH_CFG(PHYS, T, P, H, NG, R, MINB, false, true)>::launch,      \
         &::exahype::Fv3dMarchLauncher<EXAHYPE_MARCH_CFG(PHYS, T, P, H, NG, R, MINB, true, true)>::launch},      \
        {&::exahype::Fv3dMarchLauncher<EXAHYPE_MARCH_CFG(PHYS, T, P, H, NG, R, MINB, false, false)>::prepare,    \
         &::exahype::Fv3dMarchLauncher<EXAHYPE_MARCH_CFG(PHYS, T, P, H, NG, R, MINB, true, false)>::prepare,     \
         &::exahype::Fv3dMarchLauncher<EXAHYPE_MARCH_CFG(PHYS, T, P, H, NG, R, MINB, false, true)>::prepare,     \
         &::exahype::Fv3dMarchLauncher<EXAHYPE_MARCH_CFG(PHYS, T, P, H, NG, R, MINB, true, true)>::prepare},     \
        {&::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 3, P, H, G, NT, MINB_CELL, false, false)>::launch,   \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 3, P, H, G, NT, MINB_CELL, true, false)>::launch,    \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 3, P, H, G, NT, MINB_CELL, false, true)>::launch,    \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 3, P, H, G, NT, MINB_CELL, true, true)>::launch},    \
        {&::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 3, P, H, G, NT, MINB_CELL, false, false)>::prepare,  \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 3, P, H, G, NT, MINB_CELL, true, false)>::prepare,   \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 3, P, H, G, NT, MINB_CELL, false, true)>::prepare,   \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 3, P, H, G, NT, MINB_CELL, true, true)>::prepare}    \
  }

#define EXAHYPE_MARCH2D_CFG(PHYS, T, P, H, WPC, MINB, PF, DA, UH, WHICH)                                           \
  ::exahype::Fv2dMarchConfig<PHYS, ::exahype::RusanovUpdate, T, P, H, WPC, MINB, DA, UH,                       \
                             ::exahype::Fv2dVec<T, PHYS::NR + PHYS::NA>::WHICH, PF>
#define EXAHYPE_MARCH2D_DISPATCH(PHYS, T, P, H, WPC, MINB, PF, DA, UH)                                             \
  ::exahype::Fv2dMarchDispatch<EXAHYPE_MARCH2D_CFG(PHYS, T, P, H, WPC, MINB, PF, DA, UH, WIDE),                     \
                               EXAHYPE_MARCH2D_CFG(PHYS, T, P, H, WPC, MINB, PF, DA, UH, NARROW)>

// 2-D shape served by the row-marching kernel (WPC warps per CTA, register prefetch distance PF rows), with the thread-per-cell kernel (G, NT, MINB_CELL) as
// alternative
#define EXAHYPE_FV2D_ENTRY(MODEL, DTYPE, PHYS, T, P, H, WPC, MINB, PF, G, NT, MINB_CELL)                            \
  {                                                                                                       \
    {MODEL, DTYPE, 2, P, H, PHYS::NR, PHYS::NA, 0u},                                                      \
        {&EXAHYPE_MARCH2D_DISPATCH(PHYS, T, P, H, WPC, MINB, PF, false, false)::launch,                              \
         &EXAHYPE_MARCH2D_DISPATCH(PHYS, T, P, H, WPC, MINB, PF, true, false)::launch,                               \
         &EXAHYPE_MARCH2D_DISPATCH(PHYS, T, P, H, WPC, MINB, PF, false, true)::launch,                               \
         &EXAHYPE_MARCH2D_DISPATCH(PHYS, T, P, H, WPC, MINB, PF, true, true)::launch},                               \
        {&EXAHYPE_MARCH2D_DISPATCH(PHYS, T, P, H, WPC, MINB, PF, false, false)::prepare,                             \
         &EXAHYPE_MARCH2D_DISPATCH(PHYS, T, P, H, WPC, MINB, PF, true, false)::prepare,                              \
         &EXAHYPE_MARCH2D_DISPATCH(PHYS, T, P, H, WPC, MINB, PF, false, true)::prepare,                              \
         &EXAHYPE_MARCH2D_DISPATCH(PHYS, T, P, H, WPC, MINB, PF, true, true)::prepare},                              \
        {&::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 2, P, H, G, NT, MINB_CELL, false, false)>::launch,   \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 2, P, H, G, NT, MINB_CELL, true, false)>::launch,    \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 2, P, H, G, NT, MINB_CELL, false, true)>::launch,    \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 2, P, H, G, NT, MINB_CELL, true, true)>::launch},    \
        {&::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 2, P, H, G, NT, MINB_CELL, false, false)>::prepare,  \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 2, P, H, G, NT, MINB_CELL, true, false)>::prepare,   \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 2, P, H, G, NT, MINB_CELL, false, true)>::prepare,   \
         &::exahype::FvLauncher<EXAHYPE_FV_CFG(PHYS, T, 2, P, H, G, NT, MINB_CELL, true, true)>::prepare}    \
  }

}  // namespace exahype
