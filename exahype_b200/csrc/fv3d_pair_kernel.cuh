// sm_100a warp-per-patch plane-marching kernel for the 3-D batched stateless FV Rusanov patch update, 8x8x8 patches.
//
// Same arithmetic, statement order and results as fv_patch_kernel.cuh / fv3d_march_kernel.cuh (reference
// "Unit test/test.cpp":11-104 with the loop ranges of exahype/printers/CPPPrinter.py:116-137).  The group-of-three-warps
// kernel (fv3d_march_kernel.cuh) spends 29 % of its warp samples at the group's named barrier, a third of its warps
// (the face warps) idle, and its interior threads run one dependent fp64 chain each (profiles/r01_ncu_c3_march.txt:
// stall_wait 26 %).  Here ONE WARP owns a patch and marches through its planes along axis 0:
//
//   * lane <-> two vertically adjacent columns (rows 2jp and 2jp+1, same k) plus one of the 32 face-halo columns: three
//     independent instruction chains per lane, and the axis-1 exchange between the two rows never leaves registers;
//   * the axis-0 stencil is a rolling register window {i-1, i, i+1} of state, F_0, L_0 and the per-cell primitives;
//   * the step for plane i loads plane i+1 (its F_0 / L_0 complete the axis-0 stencil of plane i), evaluates F_1, F_2,
//     L_1, L_2 of plane i, publishes what other lanes need in warp-private shared scratch, and updates plane i in the same
//     step: nothing but the window is carried from plane to plane, the scratch is single buffered;
//   * all synchronisation is __syncwarp (two per plane).  Warps are independent: no named barriers, no CTA barrier after
//     start-up, every warp has its own TMA ring (1-D bulk copies + mbarrier complete_tx), scratch and output staging;
//   * finished planes leave through SB staging buffers (one by default: the shared memory goes to a fourth ring slot
//     instead) and TMA bulk stores; the cursors of ring, stream and output are warp-uniform and advanced by all lanes,
//     so that the TMA instructions take uniform-register operands instead of sitting in single-lane divergent blocks;
//   * multi-GPU: the all-reduce(max) of lambda_max can run in this kernel's epilogue (peer_mail.cuh).
//
// Bank-conflict-free for fp64: a half-warp holds row pairs {0, 2} or {1, 3} (rows 0/1 and 4/5, resp. 2/3 and 6/7), the
// AoS plane has a cell stride of 5 doubles, scratch rows are pitched 10, the staging buffer is two padded segments.
#pragma once

#include <cstdlib>

#include "fv3d_march_kernel.cuh"

namespace exahype {

template <class Upd, typename T, class = void>
struct has_dissipation_m : std::false_type {};
template <class Upd, typename T>
struct has_dissipation_m<Upd, T, std::void_t<decltype(&Upd::template dissipation_m<T>)>> : std::true_type {};

__device__ __forceinline__ void tma_store_wait_read_all_but_one() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

template <class Phys_, class Upd_, typename T_, int P_, int H_, int NW_, int R_, bool DISS_ALL_, bool UNHALOED_,
          bool GATHER_ = false, int SB_ = 2>
struct Fv3dPairConfig {
  using Phys = Phys_;
  using Upd = Upd_;
  using T = T_;
  static constexpr int DIM = 3, P = P_, H = H_, NW = NW_, R = R_, SB = SB_;   // SB: output staging buffers per warp
  static constexpr bool DISS_ALL = DISS_ALL_, UNHALOED = UNHALOED_, GATHER = GATHER_;
  static_assert(P == 8 && H == 1, "one warp per 8x8 plane: 32 row pairs and 32 face-halo columns");
  static_assert(NW >= 1 && NW <= 32 && R >= 3 && (SB == 1 || SB == 2), "pair-march geometry");

  static constexpr int NR = Phys::NR, NA = Phys::NA, NV = NR + NA;
  static constexpr int S = P + 2 * H;
  static constexpr int NPL = P + 2;
  static constexpr int PLANE_ELEMS = S * S * NV;
  static constexpr int PLANE_BYTES = PLANE_ELEMS * (int)sizeof(T);
  static constexpr int PATCH_ELEMS = S * PLANE_ELEMS;
  static constexpr int OUT_PLANE_ELEMS = P * P * NV;
  static constexpr int OUT_PATCH_ELEMS = P * OUT_PLANE_ELEMS;
  static_assert(PLANE_BYTES % 16 == 0, "plane must be a whole number of 16-byte units for TMA bulk copies");
  // the two halo planes of a patch are only read at interior (j, k): rows H .. H+P-1 are one contiguous run
  static constexpr int ROW_BYTES = S * NV * (int)sizeof(T);
  static constexpr bool TRIM_HALO_PLANES = (ROW_BYTES % 16 == 0);
  static constexpr int HALO_PLANE_SKIP_ELEMS = TRIM_HALO_PLANES ? H * S * NV : 0;
  static constexpr int HALO_PLANE_BYTES = TRIM_HALO_PLANES ? P * ROW_BYTES : PLANE_BYTES;

  static constexpr int NT = NW * 32;
  static constexpr int DV = DISS_ALL ? NR : 1;

  static constexpr int PJ = 10;                     // F_1 scratch: [x_j in 0..P+1][k], pitch PJ
  static constexpr int PK = P + 2;                  // F_2 scratch: [j][x_k in 0..P+1], pitch PK
  static constexpr int SJ = (P + 2) * PJ;
  static constexpr int SK = P * PK;
  static constexpr int SEG_ELEMS = OUT_PLANE_ELEMS / 2;                 // staging: rows j < 4 | j >= 4
  static constexpr int SEG_PITCH = SEG_ELEMS + 64 / (int)sizeof(T);     // +64 bytes
  static constexpr int STAGE_ELEMS = 2 * SEG_PITCH;
  static constexpr bool USE_TMA_STORE = UNHALOED && ((SEG_ELEMS * (int)sizeof(T)) % 16 == 0) &&
                                        ((SEG_PITCH * (int)sizeof(T)) % 16 == 0);

  // Scratch layout of F_1 / L_1 and F_2 / L_2: component planes F[v][slot], L[slot], 64-bit accesses (default).
  // -DEXAHYPE_3D_REC_SCRATCH=1 (fp64, even record width) keeps ONE record {F[0..NR), L} of RW values per slot instead
  // and moves it with 128-bit accesses: a third fewer shared-memory instructions (212 instead of 311 in the kernel), same
  // bytes, same bits -- and measured SLOWER in every regime (C3 burst 0.3069 -> 0.3140 ms, sustained 0.351 -> 0.357, fast
  // arithmetic 0.2944 -> 0.3005, profiles/r02_c3_variants.txt): the kernel is not issue-bound on its LDS/STS, and a
  // 48-byte record pitch costs the wide accesses more wavefronts than the instructions it saves.  Kept as a tuning switch.
  static constexpr int RW = NR + 1;
#ifndef EXAHYPE_3D_REC_SCRATCH
#define EXAHYPE_3D_REC_SCRATCH 0
#endif
  static constexpr bool REC = EXAHYPE_3D_REC_SCRATCH && sizeof(T) == 8 && (RW % 2 == 0);
  // per-warp shared memory
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_FJ = align_up(OFF_RING + R * PLANE_BYTES, 16);                      // REC: Rj [SJ][RW]
  static constexpr int OFF_FK = align_up(OFF_FJ + (REC ? RW : NR) * SJ * (int)sizeof(T), 16);  // REC: Rk [SK][RW]
  static constexpr int OFF_LJ = align_up(OFF_FK + (REC ? RW : NR) * SK * (int)sizeof(T), 16);
  static constexpr int OFF_LK = align_up(OFF_LJ + (REC ? 0 : SJ) * (int)sizeof(T), 16);
  static constexpr int OFF_STAGE = align_up(OFF_LK + (REC ? 0 : SK) * (int)sizeof(T), 128);
  static constexpr int OFF_BAR = align_up(OFF_STAGE + SB * STAGE_ELEMS * (int)sizeof(T), 16);
  static constexpr int WARP_BYTES = align_up(OFF_BAR + R * 8, 128);
  static constexpr int SMEM_BYTES = NW * WARP_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "warps do not fit the 227 KB of shared memory per CTA");

  static __device__ __forceinline__ int stage_index(int j, int k) {     // element offset of cell (j,k), variable 0
    return (j >> 2) * SEG_PITCH + ((j & 3) * P + k) * NV;
  }
};

// Per-lane view of one warp's stream of planes: plane n of the stream lives in ring slot n % R.
template <class C>
struct PairStream {
  using T = typename C::T;
  const T* q_in;
  T* q_out;
  T* lambda_patch;
  T *ring, *Fj, *Fk, *Lj, *Lk, *stage;
  unsigned long long* full;
  long long n_warps;  // patches between two consecutive patches of this warp
  T dt;
  int n_my_patches;
  int lane;
  // Every cursor below is warp-uniform and advanced by all lanes, so that it can live in uniform registers; only the
  // TMA / mbarrier instructions themselves are issued by lane 0.
  // consumer: patch counter, patch index, output base, ring slot (+ mbarrier phase parity) of the next plane
  int pi, slot;
  long long patch;
  T* out_base;
  uint32_t parity;
  // producer: the next plane to request -- plane p_ip of the patch at p_src, into ring slot p_slot; p_left planes to go
  const T* p_src;
  long long p_patch;
  int p_ip, p_slot, p_left;
  // CellData form: the table entries (QIn, QOut, dt) of this warp's NEXT patch are read one patch ahead, so that the
  // dependent loads -- pointer first, then the data behind it -- are not waited for when the stream gets there
  const T* p_src_next;
  T* out_next;
  T dt_next;

  __device__ __forceinline__ void issue_next_load(const FvGather<T>& gather) {
    if (p_left == 0) return;
    const bool halo_plane = (p_ip == 0) || (p_ip == C::NPL - 1);
    const int skip = halo_plane ? C::HALO_PLANE_SKIP_ELEMS : 0;
    const uint32_t bytes = halo_plane ? C::HALO_PLANE_BYTES : C::PLANE_BYTES;
    if (lane == 0) {
      mbar_expect_tx(&full[p_slot], bytes);
      tma_load_1d(ring + p_slot * C::PLANE_ELEMS + skip, p_src + (p_ip + C::H - 1) * C::PLANE_ELEMS + skip, bytes,
                  &full[p_slot]);
    }
    --p_left;
    if (++p_ip == C::NPL) {
      p_ip = 0;
      p_patch += n_warps;
      if (C::GATHER) {
        p_src = p_src_next;                                                 // valid whenever p_left > 0
        if (p_left > C::NPL) p_src_next = gather.q_in[p_patch + n_warps];   // the patch after the one just begun
      } else p_src += n_warps * C::PATCH_ELEMS;
    }
    if (++p_slot == C::R) p_slot = 0;
  }
  // consumer: start of the next patch of this warp (pi already advanced; first == true for the first one)
  __device__ __forceinline__ void begin_patch(const FvGather<T>& gather, bool first) {
    if (!first) patch += n_warps;
    constexpr int ELEMS = C::UNHALOED ? C::OUT_PATCH_ELEMS : C::PATCH_ELEMS;
    if (C::GATHER) {
      if (first) {
        out_base = gather.q_out[patch];
        if (gather.dt != nullptr) dt = gather.dt[patch];    // CellData::dt of this patch
      } else {
        out_base = out_next;
        if (gather.dt != nullptr) dt = dt_next;
      }
      if (pi + 1 < n_my_patches) {
        out_next = gather.q_out[patch + n_warps];
        if (gather.dt != nullptr) dt_next = gather.dt[patch + n_warps];
      }
    } else if (first) {
      out_base = q_out + patch * ELEMS;
    } else {
      out_base += n_warps * ELEMS;
    }
  }
  __device__ __forceinline__ const T* wait_plane() {
    mbar_wait(&full[slot], parity);
    return ring + slot * C::PLANE_ELEMS;
  }
  __device__ __forceinline__ const T* previous_plane() const {
    return ring + (slot == 0 ? C::R - 1 : slot - 1) * C::PLANE_ELEMS;
  }
  __device__ __forceinline__ void advance_plane() {
    if (++slot == C::R) { slot = 0; parity ^= 1u; }
  }

  // After the closing __syncwarp of a step: write out zero-based interior plane `plane` of patch pi from staging `buffer`.
  __device__ __forceinline__ void drain_staged_plane(const FvGather<T>& gather, int plane, int buffer) {
    const T* sbuf = stage + buffer * C::STAGE_ELEMS;
    if (C::UNHALOED) {
      T* dst = out_base + plane * C::OUT_PLANE_ELEMS;
      if (C::USE_TMA_STORE) {
        if (lane == 0) {
          tma_store_1d(dst, sbuf, C::SEG_ELEMS * (uint32_t)sizeof(T));
          tma_store_1d(dst + C::SEG_ELEMS, sbuf + C::SEG_PITCH, C::SEG_ELEMS * (uint32_t)sizeof(T));
          tma_store_commit();
          if (C::SB == 2) tma_store_wait_read_all_but_one();    // the other staging buffer (written next step) is free again
        }
      } else {
        for (int e = lane; e < C::OUT_PLANE_ELEMS; e += 32) {
          const int sgm = e / C::SEG_ELEMS;
          dst[e] = sbuf[sgm * C::SEG_PITCH + (e - sgm * C::SEG_ELEMS)];
        }
      }
    } else {
      // haloed layout: interior rows of the plane are runs of P*NV values (test.cpp:96-104 writes all NV)
      T* dst = out_base + (plane + C::H) * C::PLANE_ELEMS;
      constexpr int ROW = C::P * C::NV;
      constexpr int ROW_GAP = C::S * C::NV - ROW;        // the halo cells between two interior rows
      if constexpr (ROW >= 32 && C::SEG_ELEMS % 32 == 0 && C::OUT_PLANE_ELEMS % 32 == 0) {
        // Value e = lane + 32 i of the plane: with rows of at least a warp's width and segments of whole warps, its
        // staging segment is the same for every lane and its row is that of value 32 i or the next one -- every offset
        // is a compile-time constant but for one compare per store (the general loop below costs two divisions and
        // a dozen integer instructions per value: a third of the step's instructions in the reference's in-place form).
        const T* src = sbuf + lane;
        T* d0 = dst + (C::H * C::S + C::H) * C::NV + lane;
#pragma unroll
        for (int i = 0; i < C::OUT_PLANE_ELEMS / 32; ++i) {
          const int e0 = 32 * i, row0 = e0 / ROW, col0 = e0 % ROW;
          const int s_off = e0 + (e0 / C::SEG_ELEMS) * (C::SEG_PITCH - C::SEG_ELEMS);
          const int wrap = (lane + col0 >= ROW) ? ROW_GAP : 0;
          d0[e0 + row0 * ROW_GAP + wrap] = src[s_off];
        }
      } else {
        for (int e = lane; e < C::OUT_PLANE_ELEMS; e += 32) {
          const int row = e / ROW;
          const int sgm = e / C::SEG_ELEMS;
          dst[((row + C::H) * C::S + C::H) * C::NV + (e - row * ROW)] = sbuf[sgm * C::SEG_PITCH + (e - sgm * C::SEG_ELEMS)];
        }
      }
    }
  }
};

// what a lane owns: the cells (ja, k) and (ja + 1, k) of every plane and one face-halo column of axis 1 or 2
template <class C>
struct PairLane {
  using T = typename C::T;
  int cell;          // haloed in-plane cell index of (ja, k); the partner row is cell + S
  int sj, sk;        // scratch slots of (ja, k): partner row at sj + PJ / sk + PK
  int st;            // staging offsets of (ja, k); the partner row is st + P*NV (same segment)
  int f_cell, f_axis, f_comp_stride;
  T *f_F, *f_L;      // where the face column's F / L go (REC: f_F is the slot's record, f_L unused)
};

// one slot of the F / L scratch: store, load (layout per Fv3dPairConfig::REC)
template <class C>
__device__ __forceinline__ void pair_scratch_put(typename C::T* F_base, typename C::T* L_base, int slot, int comp_stride,
                                                 const typename C::T (&F)[C::NR], typename C::T L) {
  using T = typename C::T;
  if constexpr (C::REC) {
    double2* rec = reinterpret_cast<double2*>(F_base + slot * C::RW);
#pragma unroll
    for (int i = 0; i < C::RW / 2; ++i)
      rec[i] = make_double2(2 * i < C::NR ? F[2 * i < C::NR ? 2 * i : 0] : L, 2 * i + 1 < C::NR ? F[2 * i + 1 < C::NR ? 2 * i + 1 : 0] : L);
  } else {
#pragma unroll
    for (int v = 0; v < C::NR; ++v) F_base[v * comp_stride + slot] = F[v];
    L_base[slot] = L;
  }
}
template <class C>
__device__ __forceinline__ void pair_scratch_get(const typename C::T* F_base, const typename C::T* L_base, int slot,
                                                 int comp_stride, typename C::T (&F)[C::NR], typename C::T& L) {
  if constexpr (C::REC) {
    const double2* rec = reinterpret_cast<const double2*>(F_base + slot * C::RW);
    double tmp[C::RW];
#pragma unroll
    for (int i = 0; i < C::RW / 2; ++i) {
      const double2 v = rec[i];
      tmp[2 * i] = v.x;
      tmp[2 * i + 1] = v.y;
    }
#pragma unroll
    for (int v = 0; v < C::NR; ++v) F[v] = tmp[v];
    L = tmp[C::NR];
  } else {
#pragma unroll
    for (int v = 0; v < C::NR; ++v) F[v] = F_base[v * comp_stride + slot];
    L = L_base[slot];
  }
}

#ifndef EXAHYPE_3D_EARLY_HALOED
#define EXAHYPE_3D_EARLY_HALOED 1
#endif
#ifndef EXAHYPE_3D_EARLY2
#define EXAHYPE_3D_EARLY2 1
#endif

// the axis-0 window of one lane: three planes x two cells, all indices compile-time
template <class C>
struct PairWindow {
  using T = typename C::T;
  using Prims = typename C::Phys::template Prims<T>;
  T q[3][2][C::NV];
  T fi[3][2][C::NR];
  T li[3][2];
  Prims pr[3][2];
  T m0[2];     // max(L_0 of the current plane, L_0 of the one before it): computed as the previous plane's "plus" maximum
  // Functor families WITHOUT a per-cell cache (generated from the user's device source with the Functions.h signatures:
  // every Flux / maxEigenvalue call evaluates 1/rho, p, c itself) get all three axes of a plane evaluated when the plane
  // arrives -- one block of code, so the compiler shares the division and the root between the calls -- and F_1, F_2,
  // L_1, L_2 wait here for the step that updates the plane.  With a cache the two evaluations a step apart share it instead.
  static constexpr bool STASH = std::is_empty<Prims>::value;
  // EXAHYPE_3D_EARLY2 (one dissipated variable, dense batch): everything a lane reads from a plane's ring slot -- its two
  // cells, its face-halo column and the six neighbour values of the dissipation -- is read when the plane ARRIVES and
  // waits here for the step that updates the plane, so the slot goes back to the TMA a whole step earlier: a fifth plane
  // in flight per warp without the shared memory of a fifth slot.  254 registers (206 without), no spills; C3 0.3066 ->
  // 0.3032 ms, 4 096 patches 0.0505 -> 0.0499 (profiles/r02_s3_early_slot_release.txt).  The CellData form and the
  // cache-less families have no registers left for it (they would spill) and keep releasing the slot a step later; so do
  // the haloed output form (in place 0.332 -> 0.343 ms with it) and the geometry with two staging buffers that the
  // fast arithmetic runs (0.293 -> 0.338 ms with it): measured in the same call, same file.
  static constexpr bool EARLY2 = EXAHYPE_3D_EARLY2 && (C::DV == 1) && C::UNHALOED && (C::SB == 1) && !C::GATHER && !STASH;
  T fq[EARLY2 ? 3 : 1][C::NV], qn[EARLY2 ? 3 : 1][6];
  T fjs[STASH ? 3 : 1][2][C::NR], fks[STASH ? 3 : 1][2][C::NR];
  T ljs[STASH ? 3 : 1][2], lks[STASH ? 3 : 1][2];
};

// plane `W` of the window <- the plane the stream delivers next: state, primitives, F_0, L_0
template <class C, int W, bool EXTRAS = true>
__device__ __forceinline__ void pair_load_plane(PairStream<C>& ps, const PairLane<C>& ln, PairWindow<C>& w) {
  using T = typename C::T;
  using Phys = typename C::Phys;
  const T* __restrict__ qs = ps.wait_plane();
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int v = 0; v < C::NV; ++v) w.q[W][c][v] = qs[(ln.cell + c * C::S) * C::NV + v];
  if constexpr (PairWindow<C>::EARLY2 && EXTRAS) {
    // (a halo plane arriving as the last plane of a patch is never updated: what is read of it here is not used)
#pragma unroll
    for (int v = 0; v < C::NV; ++v) w.fq[W][v] = qs[ln.f_cell * C::NV + v];
    w.qn[W][0] = qs[(ln.cell - C::S) * C::NV];
    w.qn[W][1] = qs[(ln.cell + 2 * C::S) * C::NV];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      w.qn[W][2 + 2 * c] = qs[(ln.cell + c * C::S - 1) * C::NV];
      w.qn[W][3 + 2 * c] = qs[(ln.cell + c * C::S + 1) * C::NV];
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    w.pr[W][c] = Phys::template prims<T>(w.q[W][c]);
    Phys::template flux<0, T>(w.q[W][c], w.pr[W][c], w.fi[W][c]);
    w.li[W][c] = Phys::template eigen<0, T>(w.q[W][c], w.pr[W][c]);
    if constexpr (PairWindow<C>::STASH) {
      Phys::template flux<1, T>(w.q[W][c], w.pr[W][c], w.fjs[W][c]);
      w.ljs[W][c] = Phys::template eigen<1, T>(w.q[W][c], w.pr[W][c]);
      Phys::template flux<2, T>(w.q[W][c], w.pr[W][c], w.fks[W][c]);
      w.lks[W][c] = Phys::template eigen<2, T>(w.q[W][c], w.pr[W][c]);
    }
  }
}

// the first two planes of a patch (halo plane 0, interior plane 1) only fill the window
template <class C, int W>
__device__ __forceinline__ void pair_pre_step(PairStream<C>& ps, const FvGather<typename C::T>& gather,
                                              const PairLane<C>& ln, PairWindow<C>& w) {
  pair_load_plane<C, W, W == 1>(ps, ln, w);
  // The slot requested next is that of the previous plane of the stream.  W == 0: the last plane of the previous patch,
  // last read before the closing __syncwarp of the previous step.  W == 1: plane 0, which every lane has read once the
  // warp meets here (halo planes are never read from the ring again).
  // EARLY2: the slot of the plane just read, once every lane has read it.
  if (W == 1) {
#pragma unroll
    for (int c = 0; c < 2; ++c) w.m0[c] = fv_max(w.li[1][c], w.li[0][c]);
    __syncwarp();
  } else if (PairWindow<C>::EARLY2) {
    __syncwarp();
  }
  if (PairWindow<C>::EARLY2 || W == 1 || ps.pi >= 1) ps.issue_next_load(gather);
  ps.advance_plane();
}

// The same two planes in ONE block of code (-DEXAHYPE_3D_MERGED_PRE=1): four independent cells per lane instead of two
// after each other -- the two pre-steps have no update to overlap their division / root chains with.  Measured: no gain
// (C3 0.3080 -> 0.3098 ms in a burst, 0.3717 -> 0.3719 sustained, profiles/r02_s3_merged_pre_steps.txt): the request for
// the previous patch's last slot leaves half a pre-step later, which costs what the shorter chains save.  Off.
#ifndef EXAHYPE_3D_MERGED_PRE
#define EXAHYPE_3D_MERGED_PRE 0
#endif
template <class C>
__device__ __forceinline__ void pair_pre_steps_merged(PairStream<C>& ps, const FvGather<typename C::T>& gather,
                                                      const PairLane<C>& ln, PairWindow<C>& w) {
  using T = typename C::T;
  using Phys = typename C::Phys;
  static_assert(!PairWindow<C>::EARLY2, "the merged pre-steps keep the one-step-later slot release");
  const T* __restrict__ q0s = ps.wait_plane();
  ps.advance_plane();
  const T* __restrict__ q1s = ps.wait_plane();
  ps.advance_plane();
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int v = 0; v < C::NV; ++v) {
      w.q[0][c][v] = q0s[(ln.cell + c * C::S) * C::NV + v];
      w.q[1][c][v] = q1s[(ln.cell + c * C::S) * C::NV + v];
    }
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int W = 0; W < 2; ++W) {
      w.pr[W][c] = Phys::template prims<T>(w.q[W][c]);
      Phys::template flux<0, T>(w.q[W][c], w.pr[W][c], w.fi[W][c]);
      w.li[W][c] = Phys::template eigen<0, T>(w.q[W][c], w.pr[W][c]);
    }
  if constexpr (PairWindow<C>::STASH) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      Phys::template flux<1, T>(w.q[1][c], w.pr[1][c], w.fjs[1][c]);
      w.ljs[1][c] = Phys::template eigen<1, T>(w.q[1][c], w.pr[1][c]);
      Phys::template flux<2, T>(w.q[1][c], w.pr[1][c], w.fks[1][c]);
      w.lks[1][c] = Phys::template eigen<2, T>(w.q[1][c], w.pr[1][c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) w.m0[c] = fv_max(w.li[1][c], w.li[0][c]);
  __syncwarp();
  // the two slots requested now: the last plane of the previous patch (none before the first patch: the ring was filled
  // at start-up) and plane 0, both read by every lane before the __syncwarp
  if (ps.pi >= 1) ps.issue_next_load(gather);
  ps.issue_next_load(gather);
}

// Interior plane ip (1..P) of the current patch, window phase PH = ip % 3:
//   plane ip+1 -> window;  F_1, F_2, L_1, L_2 of plane ip (+ this lane's face column) -> registers / scratch;  __syncwarp;
//   update plane ip -> staging;  __syncwarp;  drain;  request the next plane of the stream.
template <class C, int PH>
__device__ __forceinline__ void pair_main_step(PairStream<C>& ps, const FvGather<typename C::T>& gather, int ip,
                                               const PairLane<C>& ln, PairWindow<C>& w, typename C::T& lam_local,
                                               typename C::T& warp_lam) {
  using T = typename C::T;
  using Phys = typename C::Phys;
  using Upd = typename C::Upd;
  constexpr int NV = C::NV, NR = C::NR, SJ = C::SJ, SK = C::SK, PJ = C::PJ, PK = C::PK, S = C::S;
  constexpr int NEW = (PH + 1) % 3, MID = PH, OLD = (PH + 2) % 3;
  const int wb = (C::SB == 2) ? (ip & 1) : 0;

  pair_load_plane<C, NEW>(ps, ln, w);
  const T* __restrict__ qm = ps.previous_plane();        // plane ip in the ring: face columns, neighbours' Q

  // ------------------------------------------------------------ plane ip: F_1, F_2, L_1, L_2
  T fj[2][NR], lj[2], lk[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    T F[NR];
    if constexpr (PairWindow<C>::STASH) {
#pragma unroll
      for (int v = 0; v < NR; ++v) { fj[c][v] = w.fjs[MID][c][v]; F[v] = w.fks[MID][c][v]; }
      lj[c] = w.ljs[MID][c];
      lk[c] = w.lks[MID][c];
    } else {
      Phys::template flux<1, T>(w.q[MID][c], w.pr[MID][c], fj[c]);
      lj[c] = Phys::template eigen<1, T>(w.q[MID][c], w.pr[MID][c]);
      Phys::template flux<2, T>(w.q[MID][c], w.pr[MID][c], F);
      lk[c] = Phys::template eigen<2, T>(w.q[MID][c], w.pr[MID][c]);
    }
    pair_scratch_put<C>(ps.Fk, ps.Lk, ln.sk + c * PK, SK, F, lk[c]);
    // the row above (c = 0) / below (c = 1) belongs to another lane: it reads this row's F_1 / L_1 from the scratch
    pair_scratch_put<C>(ps.Fj, ps.Lj, ln.sj + c * PJ, SJ, fj[c], lj[c]);
    lam_local = fv_max(lam_local, fv_max(w.li[MID][c], fv_max(lj[c], lk[c])));
  }
  // this lane's face-halo column: F_axis / L_axis of the cell one layer outside the interior (one instruction stream for
  // both axes, see march_face_eval)
  {
    T q[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) q[v] = PairWindow<C>::EARLY2 ? w.fq[MID][v] : qm[ln.f_cell * NV + v];
    const auto pr = Phys::template prims<T>(q);
    T F[NR];
    T L;
    if constexpr (has_runtime_axis<Phys, T>::value) {
      Phys::template flux_runtime<T>(q, pr, ln.f_axis, F);
      L = Phys::template eigen_runtime<T>(q, pr, ln.f_axis);
    } else if (ln.f_axis == 1) {
      Phys::template flux<1, T>(q, pr, F);
      L = Phys::template eigen<1, T>(q, pr);
    } else {
      Phys::template flux<2, T>(q, pr, F);
      L = Phys::template eigen<2, T>(q, pr);
    }
    pair_scratch_put<C>(ln.f_F, ln.f_L, 0, ln.f_comp_stride, F, L);
  }
  // With one dissipated variable the neighbours' Q the dissipation needs from plane ip are few: read them now, so that
  // nothing touches the ring slot of plane ip after this point and the next plane of the stream can be requested right
  // behind the __syncwarp instead of at the end of the step (half a step more lead for the TMA).
#ifndef EXAHYPE_3D_INTERLEAVE
#define EXAHYPE_3D_INTERLEAVE 1
#endif
  constexpr bool EARLY = (C::DV == 1) && (C::UNHALOED || EXAHYPE_3D_EARLY_HALOED);
  T qn_j[2], qn_k[2][2];   // [cell]: the neighbour across axis 1 outside the pair; [cell][-1 / +1] along axis 2
  if constexpr (PairWindow<C>::EARLY2) {
    qn_j[0] = w.qn[MID][0];
    qn_j[1] = w.qn[MID][1];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      qn_k[c][0] = w.qn[MID][2 + 2 * c];
      qn_k[c][1] = w.qn[MID][3 + 2 * c];
    }
  } else if constexpr (EARLY) {
    qn_j[0] = qm[(ln.cell - S) * NV];
    qn_j[1] = qm[(ln.cell + 2 * S) * NV];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      qn_k[c][0] = qm[(ln.cell + c * S - 1) * NV];
      qn_k[c][1] = qm[(ln.cell + c * S + 1) * NV];
    }
  }
  if (C::SB == 1 && C::USE_TMA_STORE && ps.lane == 0) tma_store_wait_read();   // the one staging buffer is free again
  __syncwarp();
  if constexpr (EARLY) ps.issue_next_load(gather);

  // ------------------------------------------------------------ update plane ip
  const T dt = ps.dt;
  constexpr bool SHARE_MAX = has_dissipation_m<Upd, T>::value;
  T mj = T(0);
  if constexpr (SHARE_MAX) mj = fv_max(lj[1], lj[0]);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int cell = ln.cell + c * S;
    // the neighbours' F / L: across axis 1 the row outside the pair (above for c = 0, below for c = 1), along axis 2 both sides
    T Fjn[NR], Ljn, Fkp[NR], Lkp, Fkm[NR], Lkm;
    pair_scratch_get<C>(ps.Fj, ps.Lj, ln.sj + (c == 0 ? -PJ : 2 * PJ), SJ, Fjn, Ljn);
    pair_scratch_get<C>(ps.Fk, ps.Lk, ln.sk + c * PK + 1, SK, Fkp, Lkp);
    pair_scratch_get<C>(ps.Fk, ps.Lk, ln.sk + c * PK - 1, SK, Fkm, Lkm);
    T qc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) qc[v] = w.q[MID][c][v];
    // "Q_copy = Q_copy - 0.5*F[+1] + 0.5*F[-1]" for axis 0, 1, 2 in order (test.cpp:60-77)
#pragma unroll
    for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], w.fi[NEW][c][v], w.fi[OLD][c][v]);
#pragma unroll
    for (int v = 0; v < NR; ++v)
      qc[v] = (c == 0) ? Upd::flux(qc[v], fj[1][v], Fjn[v]) : Upd::flux(qc[v], Fjn[v], fj[0][v]);
#pragma unroll
    for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], Fkp[v], Fkm[v]);
    // "Q_copy = 0.5*dt*(...) + Q_copy" from the original Q, axis 0, 1, 2 in order (test.cpp:78-95)
    if constexpr (SHARE_MAX) {
      // max(L, L') of a pair of cells serves both: along axis 0 it is carried from the previous step (m0), inside the
      // lane's row pair it is mj
      const T m_up = fv_max(w.li[NEW][c], w.li[MID][c]);
#pragma unroll
      for (int v = 0; v < C::DV; ++v)
        qc[v] = Upd::dissipation_m(qc[v], w.q[MID][c][v], w.q[NEW][c][v], w.q[OLD][c][v], m_up, w.m0[c], dt);
      w.m0[c] = m_up;
      if (c == 0) {
        const T m_minus = fv_max(Ljn, lj[0]);
#pragma unroll
        for (int v = 0; v < C::DV; ++v)
          qc[v] = Upd::dissipation_m(qc[v], w.q[MID][0][v], w.q[MID][1][v], EARLY ? qn_j[0] : qm[(cell - S) * NV + v], mj,
                                     m_minus, dt);
      } else {
        const T m_plus = fv_max(Ljn, lj[1]);
#pragma unroll
        for (int v = 0; v < C::DV; ++v)
          qc[v] = Upd::dissipation_m(qc[v], w.q[MID][1][v], EARLY ? qn_j[1] : qm[(cell + S) * NV + v], w.q[MID][0][v],
                                     m_plus, mj, dt);
      }
    } else {
#pragma unroll
      for (int v = 0; v < C::DV; ++v)
        qc[v] = Upd::dissipation(qc[v], w.q[MID][c][v], w.q[NEW][c][v], w.q[OLD][c][v], w.li[MID][c], w.li[NEW][c],
                                 w.li[OLD][c], dt);
      if (c == 0) {
        const T l_minus = Ljn;
#pragma unroll
        for (int v = 0; v < C::DV; ++v)
          qc[v] = Upd::dissipation(qc[v], w.q[MID][0][v], w.q[MID][1][v], EARLY ? qn_j[0] : qm[(cell - S) * NV + v], lj[0],
                                   lj[1], l_minus, dt);
      } else {
        const T l_plus = Ljn;
#pragma unroll
        for (int v = 0; v < C::DV; ++v)
          qc[v] = Upd::dissipation(qc[v], w.q[MID][1][v], EARLY ? qn_j[1] : qm[(cell + S) * NV + v], w.q[MID][0][v], lj[1],
                                   l_plus, lj[0], dt);
      }
    }
    {
      const T l_plus = Lkp, l_minus = Lkm;
#pragma unroll
      for (int v = 0; v < C::DV; ++v)
        qc[v] = Upd::dissipation(qc[v], w.q[MID][c][v], EARLY ? qn_k[c][1] : qm[(cell + 1) * NV + v],
                                 EARLY ? qn_k[c][0] : qm[(cell - 1) * NV + v], lk[c], l_plus, l_minus, dt);
    }
    fv_apply_source<Phys, Upd, T>(qc, w.q[MID][c], dt);       // "Q_copy = Q_copy + dt*S" (families with a source term)
    T* dst = ps.stage + wb * C::STAGE_ELEMS + ln.st + c * (C::P * NV);
#pragma unroll
    for (int v = 0; v < NV; ++v) dst[v] = qc[v];
  }
  if (C::USE_TMA_STORE) fence_proxy_async_smem();
  // per-patch maximum eigenvalue over interior cells of the input state: complete at the patch's last interior plane
  if (ip == C::P) {
    T m = lam_local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fv_max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (ps.lane == 0 && ps.lambda_patch) ps.lambda_patch[ps.patch] = m;
    warp_lam = fv_max(warp_lam, m);
    lam_local = T(0);
  }
  __syncwarp();

  // ------------------------------------------------------------ drain, request the next plane
  ps.drain_staged_plane(gather, ip - 1, wb);
  // plane ip of the ring was last read by the update above: its slot takes the next plane to request
  if constexpr (!EARLY) ps.issue_next_load(gather);
  ps.advance_plane();
}

// CTAs per SM the register allocation aims at.  Tuning switch: with one-warp CTAs (-DEXAHYPE_3D_NW=1) 9 per SM put three
// warps on one scheduler, which caps a thread at 168 registers -- measured slower (profiles/r02_s3_one_warp_ctas.txt).
#ifndef EXAHYPE_3D_MINB
#define EXAHYPE_3D_MINB 1
#endif
template <class C>
__global__ void __launch_bounds__(C::NT, EXAHYPE_3D_MINB)
fv3d_pair_kernel(const typename C::T* q_in, typename C::T* q_out, long long n_patches, typename C::T dt,
                 typename C::T* __restrict__ lambda_patch, typename C::T* __restrict__ lambda_max,
                 const FvGather<typename C::T> gather) {
  using T = typename C::T;
  using Bits = typename FloatBits<T>::type;
  constexpr int P = C::P, H = C::H, S = C::S, NR = C::NR, R = C::R, NPL = C::NPL;

  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // tells the compiler that it is warp-uniform
  const int lane = threadIdx.x & 31;
  unsigned char* const ws = smem + warp * C::WARP_BYTES;

  PairStream<C> ps;
  ps.q_in = q_in; ps.q_out = q_out; ps.lambda_patch = lambda_patch; ps.dt = dt;
  ps.ring = reinterpret_cast<T*>(ws + C::OFF_RING);
  ps.Fj = reinterpret_cast<T*>(ws + C::OFF_FJ);          // [NR][SJ]   (REC: records [SJ][RW], Lj / Lk inside)
  ps.Fk = reinterpret_cast<T*>(ws + C::OFF_FK);          // [NR][SK]
  ps.Lj = reinterpret_cast<T*>(ws + C::OFF_LJ);          // [SJ]
  ps.Lk = reinterpret_cast<T*>(ws + C::OFF_LK);          // [SK]
  ps.stage = reinterpret_cast<T*>(ws + C::OFF_STAGE);    // [SB][STAGE_ELEMS]
  ps.full = reinterpret_cast<unsigned long long*>(ws + C::OFF_BAR);   // [R]
  ps.lane = lane;

  // Programmatic dependent launch (time-loop launches carry the attribute, see the launcher): let the NEXT launch of the
  // stream be scheduled right away -- its CTAs take an SM as soon as one of ours leaves and run their start-up (barrier
  // set-up, lane geometry) while the rest of this grid finishes -- ...
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (lane == 0) {
    for (int s = 0; s < R; ++s) mbar_init(&ps.full[s], 1);
    fence_mbar_init();
  }
  __syncwarp();

  ps.n_warps = (long long)gridDim.x * C::NW;
  // Patches are dealt to the warps CTA-interleaved (warp w of CTA c is worker w * grid + c): a round that does not fill
  // the grid -- the last one of a batch, or the only one of a small batch -- then spreads over all SMs with fewer warps
  // each, instead of filling some SMs and leaving the others idle (C5's small batches; EXAHYPE_3D_INTERLEAVE=0: the
  // CTA-major order of round 1).
#if EXAHYPE_3D_INTERLEAVE
  const long long w_index = (long long)warp * gridDim.x + blockIdx.x;
#else
  const long long w_index = (long long)blockIdx.x * C::NW + warp;
#endif
  const long long my_patches = (n_patches > w_index) ? (n_patches - w_index + ps.n_warps - 1) / ps.n_warps : 0;
  ps.n_my_patches = (int)my_patches;
  ps.pi = ps.slot = 0;
  ps.parity = 0;
  ps.patch = ps.p_patch = w_index;
  ps.out_base = q_out;
  ps.p_left = (int)(my_patches * NPL);
  ps.p_ip = ps.p_slot = 0;
  ps.p_src = q_in;
  if (my_patches > 0) ps.p_src = gather.template in<C::GATHER>(q_in, w_index, C::PATCH_ELEMS);
  ps.p_src_next = q_in; ps.out_next = q_out; ps.dt_next = dt;
  if (C::GATHER && my_patches > 1) ps.p_src_next = gather.q_in[w_index + ps.n_warps];

  // lane -> (row pair jp, column k): a half-warp holds the pairs {0, 2} or {1, 3}
  PairLane<C> ln;
  {
    const int k = lane & 7;
    const int jp = 2 * ((lane >> 3) & 1) + (lane >> 4);
    const int ja = 2 * jp;
    ln.cell = (ja + H) * S + (k + H);
    ln.sj = (ja + 1) * C::PJ + k;
    ln.sk = ja * C::PK + (k + 1);
    ln.st = C::stage_index(ja, k);
    // face column f = lane -> [axis-1 low | axis-1 high | axis-2 low | axis-2 high], P columns each
    const int f_axis = (lane / (2 * P)) ? 2 : 1;
    const int f_side = (lane / P) & 1;
    const int f_pos = lane % P;
    const int edge = f_side ? H + P : H - 1;
    ln.f_axis = f_axis;
    ln.f_cell = (f_axis == 1) ? edge * S + (f_pos + H) : (f_pos + H) * S + edge;
    const int slot_in_scratch = (f_axis == 1) ? (f_side ? P + 1 : 0) * C::PJ + f_pos : f_pos * C::PK + (f_side ? P + 1 : 0);
    ln.f_F = (f_axis == 1 ? ps.Fj : ps.Fk) + slot_in_scratch * (C::REC ? C::RW : 1);
    ln.f_L = (f_axis == 1 ? ps.Lj : ps.Lk) + slot_in_scratch;
    ln.f_comp_stride = (f_axis == 1) ? C::SJ : C::SK;
  }

  PairWindow<C> w;
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int v = 0; v < C::NV; ++v) w.q[s][c][v] = T(0);
#pragma unroll
      for (int v = 0; v < NR; ++v) w.fi[s][c][v] = T(0);
      w.li[s][c] = T(0);
      w.pr[s][c] = {};
      if constexpr (PairWindow<C>::STASH) {
#pragma unroll
        for (int v = 0; v < NR; ++v) w.fjs[s][c][v] = w.fks[s][c][v] = T(0);
        w.ljs[s][c] = w.lks[s][c] = T(0);
      }
    }
  if constexpr (PairWindow<C>::EARLY2) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
#pragma unroll
      for (int v = 0; v < C::NV; ++v) w.fq[s][v] = T(0);
#pragma unroll
      for (int v = 0; v < 6; ++v) w.qn[s][v] = T(0);
    }
  }
  T lam_local = T(0), warp_lam = T(0);

  // ... and touch nothing the PREVIOUS launch may still be writing (its output may be this launch's input, its last warp
  // publishes the time step) before that grid has completed and flushed.  No-ops without the launch attribute.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int s = 0; s < R; ++s) ps.issue_next_load(gather);
  const bool first_warp = (w_index == 0);
  if (first_warp && lane == 0 && gather.peer.mode == FV_PEER_LOOP) peer_trace(gather.peer, gather.peer.seq, FV_TRACE_KERNEL_BEGIN);
  // Device-resident time step (peer_mail.cuh): the ring is requested, the lane geometry is set up; a wait for the
  // slowest peer's maximum of the previous step overlaps this launch's start-up and the first planes' flight from HBM
  // instead of having extended the previous launch.
  if (ps.n_my_patches > 0) ps.dt = peer_loop_dt<T>(gather.peer, lane, ps.dt, first_warp);
  for (; ps.pi < ps.n_my_patches; ++ps.pi) {
    ps.begin_patch(gather, ps.pi == 0);
#if EXAHYPE_3D_MERGED_PRE
    pair_pre_steps_merged<C>(ps, gather, ln, w);
#else
    pair_pre_step<C, 0>(ps, gather, ln, w);
    pair_pre_step<C, 1>(ps, gather, ln, w);
#endif
    int ip = 1;
    while (true) {
      pair_main_step<C, 1>(ps, gather, ip, ln, w, lam_local, warp_lam);
      if (++ip > P) break;
      pair_main_step<C, 2>(ps, gather, ip, ln, w, lam_local, warp_lam);
      if (++ip > P) break;
      pair_main_step<C, 0>(ps, gather, ip, ln, w, lam_local, warp_lam);
      if (++ip > P) break;
    }
  }
  if (lane == 0) {
    if (C::USE_TMA_STORE) tma_store_wait_all();
    if (lambda_max != nullptr && ps.n_my_patches > 0)
      atomicMax(reinterpret_cast<Bits*>(lambda_max), FloatBits<T>::to(warp_lam));
  }
  // multi-GPU: the global admissible-time-step scalar in the same launch -- the last warp of the grid to get here
  // exchanges this device's maximum with every peer over NVLink (peer_mail.cuh) and leaves the result in *lambda_max
  // (blocking), or only publishes it for the next launch's warps to consume (time loop)
  if (gather.peer.mode == FV_PEER_BLOCKING || gather.peer.mode == FV_PEER_LOOP)
    fused_allreduce_max<T, Bits>(gather.peer, lambda_max, lane, gridDim.x * (unsigned)C::NW);
}

// EXAHYPE_PDL=0 in the environment turns programmatic dependent launch off (A/B measurements)
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("EXAHYPE_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <class C>
struct Fv3dPairLauncher {
  static cudaError_t prepare(FvLaunchInfo* info, long long n_patches) {
    static int cached_ctas_per_sm[64];
    static int cached_sms[64];
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (cached_ctas_per_sm[dev] == 0) {
      err = cudaFuncSetAttribute(fv3d_pair_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
      if (err != cudaSuccess) return err;
      int per_sm = 0, sms = 0;
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fv3d_pair_kernel<C>, C::NT, C::SMEM_BYTES);
      if (err != cudaSuccess) return err;
      err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (err != cudaSuccess) return err;
      if (per_sm < 1) return cudaErrorLaunchOutOfResources;
      cached_sms[dev] = sms;
      cached_ctas_per_sm[dev] = per_sm;
    }
    const long long ctas_needed = EXAHYPE_3D_INTERLEAVE ? n_patches : (n_patches + C::NW - 1) / C::NW;
    const long long resident = (long long)cached_sms[dev] * cached_ctas_per_sm[dev];
    info->grid = (int)(ctas_needed < resident ? ctas_needed : resident);
    info->block = C::NT;
    info->smem_bytes = C::SMEM_BYTES;
    info->patches_per_tile = C::NW;
    info->ctas_per_sm = cached_ctas_per_sm[dev];
    info->fused_allreduce = 1;
    return cudaSuccess;
  }

  static cudaError_t launch(const void* q_in, void* q_out, long long n_patches, double dt, void* lambda_patch,
                            void* lambda_max, cudaStream_t stream, const FvGatherRaw* gather = nullptr) {
    using T = typename C::T;
    if (n_patches <= 0) return cudaSuccess;
    FvLaunchInfo info;
    cudaError_t err = prepare(&info, n_patches);
    if (err != cudaSuccess) return err;
    const FvGather<T> g = make_gather<T>(gather);
    if (g.peer.mode == FV_PEER_LOOP && pdl_enabled()) {
      // steps of a device-resident time loop follow each other with nothing in between (no memset, no second kernel):
      // programmatic stream serialization overlaps this launch's scheduling and start-up with the previous step's tail
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)info.grid);
      cfg.blockDim = dim3((unsigned)info.block);
      cfg.dynamicSmemBytes = (size_t)info.smem_bytes;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      return cudaLaunchKernelEx(&cfg, fv3d_pair_kernel<C>, static_cast<const T*>(q_in), static_cast<T*>(q_out),
                                (long long)n_patches, static_cast<T>(dt), static_cast<T*>(lambda_patch),
                                static_cast<T*>(lambda_max), g);
    }
    fv3d_pair_kernel<C><<<info.grid, info.block, info.smem_bytes, stream>>>(
        static_cast<const T*>(q_in), static_cast<T*>(q_out), n_patches, static_cast<T>(dt),
        static_cast<T*>(lambda_patch), static_cast<T*>(lambda_max), g);
    return cudaGetLastError();
  }
};

// what a generated unit (exahype.printers.CUDAPrinter) instantiates: 8 warps per CTA (registers are allocated to a CTA in
// units of four warps; fp64 needs up to 255 registers per thread, fp32 fits two such CTAs per SM) with a 4-deep ring
// and one staging buffer per warp if that fits 227 KB of shared memory, else a 3-deep ring, else fewer warps
template <class Phys, class Upd, typename T, int P, int H, bool DA, bool UH>
struct Fv3dPairAutoConfig {
  using One4 = Fv3dPairConfig<Phys, Upd, T, P, H, 1, 4, DA, UH, false, 1>;
  using One3 = Fv3dPairConfig<Phys, Upd, T, P, H, 1, 3, DA, UH, false, 1>;
  static constexpr bool DEEP = 8 * One4::WARP_BYTES + 16 <= 227 * 1024;
  static constexpr int R = DEEP ? 4 : 3;
  static constexpr int BY_SMEM = (227 * 1024 - 16) / (DEEP ? One4::WARP_BYTES : One3::WARP_BYTES);
  static constexpr int NW0 = BY_SMEM < 8 ? BY_SMEM : 8;
  static constexpr int NW = NW0 >= 4 ? NW0 / 4 * 4 : (NW0 < 1 ? 1 : NW0);
  using type = Fv3dPairConfig<Phys, Upd, T, P, H, NW, R, DA, UH, false, 1>;
  using gather_type = Fv3dPairConfig<Phys, Upd, T, P, H, NW, R, DA, UH, true, 1>;     // CellData form
};
template <class Phys, class Upd, typename T, int P, int H, bool DA, bool UH>
using Fv3dPairAuto = Fv3dPairLauncher<typename Fv3dPairAutoConfig<Phys, Upd, T, P, H, DA, UH>::type>;

}  // namespace exahype
