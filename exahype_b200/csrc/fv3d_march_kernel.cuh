// sm_100a plane-marching kernel for the 3-D batched stateless FV Rusanov patch update.
//
// Same arithmetic, statement order and results as fv_patch_kernel.cuh (reference "Unit test/test.cpp":11-104 with the
// loop ranges of exahype/printers/CPPPrinter.py:116-137), different data flow.  The thread-per-cell kernel exchanges
// F_n / L_n of all three axes through shared memory and is bound by shared-memory wavefronts (profiles/r01_*: 74 % L1
// data pipe, 37 % DRAM).  Here a *group* of warps owns one patch at a time and marches through its planes along axis 0
// (the slowest index `i`):
//
//   * thread <-> column (j,k).  The axis-0 stencil lives in registers: a rolling window {i-1, i, i+1} of the cell
//     state, F_0 and L_0.  Nothing of axis 0 ever touches shared memory.
//   * only the current plane's F_1, F_2, L_1, L_2 go through shared scratch (two alternating plane buffers, one
//     named barrier per plane, 3 warps wide for 8x8x8 patches);  face-halo columns of axes 1 and 2 are evaluated by
//     the group's last warp(s).
//   * planes stream HBM -> shared memory through a ring of R plane buffers filled by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), issued R-2 planes ahead and running seamlessly across patch boundaries;
//     the footprint of a group is ~45 KB instead of ~190 KB, so 5 independent groups are resident per SM and their
//     load / FP64 / shared-memory phases overlap instead of meeting at CTA-wide barriers.
//   * finished planes leave through a two-deep staging buffer: one TMA bulk store per plane (un-haloed output) or
//     coalesced row stores (haloed output).
//   * bank-conflict-free by construction for 8x8x8 fp64: lanes are laid out so that each half-warp holds rows j and
//     j+4 (AoS plane reads, stride 5 doubles), scratch rows are pitched 10, the staging buffer is split in two padded
//     segments.
#pragma once

#include <type_traits>

#include "fv_patch_kernel.cuh"

namespace exahype {

template <class Phys_, class Upd_, typename T_, int P_, int H_, int NG_, int R_, int MINB_, bool DISS_ALL_,
          bool UNHALOED_, bool GATHER_ = false>
struct Fv3dMarchConfig {
  using Phys = Phys_;
  using Upd = Upd_;
  using T = T_;
  static constexpr int DIM = 3, P = P_, H = H_, NG = NG_, R = R_, MINB = MINB_;
  static constexpr bool DISS_ALL = DISS_ALL_, UNHALOED = UNHALOED_, GATHER = GATHER_;
  static_assert(P >= 1 && H >= 1 && NG >= 1 && NG <= 15 && R >= 3, "march geometry");

  static constexpr int NR = Phys::NR, NA = Phys::NA, NV = NR + NA;
  static constexpr int S = P + 2 * H;
  static constexpr int NPL = P + 2;                       // planes a patch needs: one halo layer each side
  static constexpr int PLANE_CELLS = S * S;
  static constexpr int PLANE_ELEMS = PLANE_CELLS * NV;    // one full haloed plane i = const (contiguous in the AoS batch)
  static constexpr int PLANE_BYTES = PLANE_ELEMS * (int)sizeof(T);
  static constexpr int PATCH_ELEMS = S * PLANE_ELEMS;
  static constexpr int OUT_PLANE_ELEMS = P * P * NV;
  static constexpr int OUT_PATCH_ELEMS = P * OUT_PLANE_ELEMS;
  static_assert(PLANE_BYTES % 16 == 0, "plane must be a whole number of 16-byte units for TMA bulk copies");
  // the two halo planes of a patch are only read at interior (j, k): their rows H .. H+P-1 are one contiguous run, and
  // the copy skips the rest when that run keeps the 16-byte granularity of bulk copies
  static constexpr int ROW_BYTES = S * NV * (int)sizeof(T);
  static constexpr bool TRIM_HALO_PLANES = (ROW_BYTES % 16 == 0);
  static constexpr int HALO_PLANE_SKIP_ELEMS = TRIM_HALO_PLANES ? H * S * NV : 0;
  static constexpr int HALO_PLANE_BYTES = TRIM_HALO_PLANES ? P * ROW_BYTES : PLANE_BYTES;

  static constexpr int N_INT = P * P;                     // interior columns
  static constexpr int N_FACE = 4 * P;                    // face-halo columns of axes 1 and 2
  static constexpr int GROUP_THREADS = (N_INT + 31) / 32 * 32 + (N_FACE + 31) / 32 * 32;
  static constexpr int FACE_BASE = (N_INT + 31) / 32 * 32;   // first thread of the face warps
  static constexpr int NT = NG * GROUP_THREADS;
  static_assert(NT <= 1024, "too many threads per CTA");
  static constexpr int DV = DISS_ALL ? NR : 1;

  // pitches chosen so that the half-warp pairing (rows j, j+4) of 8x8 fp64 planes is bank-conflict free
  static constexpr bool PAIRED = (P == 8);
  static constexpr int PJ = PAIRED ? 10 : P;              // F_1 scratch: [x_j in 0..P+1][k], pitch PJ
  static constexpr int PK = P + 2;                        // F_2 scratch: [j][x_k in 0..P+1], pitch PK
  static constexpr int SJ = (P + 2) * PJ;
  static constexpr int SK = P * PK;
  static constexpr int STAGE_SEGS = PAIRED ? 2 : 1;       // staging buffer segments (rows j < 4 | j >= 4)
  static constexpr int SEG_ELEMS = OUT_PLANE_ELEMS / STAGE_SEGS;
  static constexpr int SEG_PITCH = SEG_ELEMS + (PAIRED ? 16 / (int)sizeof(T) * 4 : 0);   // +64 bytes
  static constexpr bool USE_TMA_STORE = UNHALOED && ((SEG_ELEMS * (int)sizeof(T)) % 16 == 0) &&
                                        ((SEG_PITCH * (int)sizeof(T)) % 16 == 0);

  // per-group shared memory
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_FJ = align_up(OFF_RING + R * PLANE_BYTES, 16);
  static constexpr int OFF_FK = align_up(OFF_FJ + 2 * NR * SJ * (int)sizeof(T), 16);
  static constexpr int OFF_LJ = align_up(OFF_FK + 2 * NR * SK * (int)sizeof(T), 16);
  static constexpr int OFF_LK = align_up(OFF_LJ + 2 * SJ * (int)sizeof(T), 16);
  static constexpr int OFF_STAGE = align_up(OFF_LK + 2 * SK * (int)sizeof(T), 128);
  static constexpr int OFF_LAM = align_up(OFF_STAGE + 2 * STAGE_SEGS * SEG_PITCH * (int)sizeof(T), 16);
  static constexpr int OFF_BAR = align_up(OFF_LAM + 2 * 8, 16);
  static constexpr int GROUP_BYTES = align_up(OFF_BAR + R * 8, 16);
  static constexpr int SMEM_BYTES = NG * GROUP_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "groups do not fit the 227 KB of shared memory per CTA");

  // column of interior thread t in [0, N_INT)
  static __device__ __forceinline__ void column(int t, int& j, int& k) {
    if (PAIRED) {           // rows in the order 0,4,1,5,2,6,3,7: every half-warp holds rows j and j+4
      k = t & 7;
      const int r = t >> 3;
      j = (r >> 1) + 4 * (r & 1);
    } else {
      j = t / P;
      k = t - j * P;
    }
  }
  static __device__ __forceinline__ int stage_index(int j, int k) {   // element offset of cell (j,k), variable 0
    if (PAIRED) return (j >> 2) * SEG_PITCH + ((j & 3) * P + k) * NV;
    return (j * P + k) * NV;
  }
};

template <class Phys, typename T, class = void>
struct has_runtime_axis : std::false_type {};
template <class Phys, typename T>
struct has_runtime_axis<Phys, T, std::void_t<decltype(&Phys::template eigen_runtime<T>)>> : std::true_type {};

__device__ __forceinline__ void named_barrier_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Per-thread view of one warp group's stream of planes.  The group consumes planes 0..P+1 of its patches back to back;
// plane n of the stream lives in ring slot n % R; within a patch, plane ip uses register-window set ip % 3 and
// scratch / staging buffer ip % 2 (all compile-time: the per-patch plane loop is fully unrolled).
template <class C>
struct MarchStream {
  using T = typename C::T;
  using Bits = typename FloatBits<T>::type;
  const T* q_in;
  T* q_out;
  T* lambda_patch;
  T *ring, *Fj, *Fk, *Lj, *Lk, *stage;
  Bits* lam_slot;
  unsigned long long* full;
  long long g_index, n_groups;
  T dt;
  int n_seq;        // planes this group streams = patches * (P+2)
  int n_my_patches;
  int gt, bar_id;
  // consumer cursor: patch counter and ring slot (+ mbarrier phase parity) of the next plane
  int pi, slot;
  uint32_t parity;
  // producer cursor (thread 0 of the group): next plane to request
  int p_seq, p_ip, p_pi, p_slot;

  // `gather` (CellData form: per-patch pointers / time steps, used by GATHER instantiations only) is always the kernel
  // parameter itself, read from the constant bank where it is used instead of occupying registers
  __device__ __forceinline__ void issue_next_load(const FvGather<T>& gather) {
    const long long patch = g_index + (long long)p_pi * n_groups;
    const bool halo_plane = (p_ip == 0) || (p_ip == C::NPL - 1);
    const int skip = halo_plane ? C::HALO_PLANE_SKIP_ELEMS : 0;
    const uint32_t bytes = halo_plane ? C::HALO_PLANE_BYTES : C::PLANE_BYTES;
    mbar_expect_tx(&full[p_slot], bytes);
    tma_load_1d(ring + p_slot * C::PLANE_ELEMS + skip,
                gather.template in<C::GATHER>(q_in, patch, C::PATCH_ELEMS) + (long long)(p_ip + C::H - 1) * C::PLANE_ELEMS + skip,
                bytes, &full[p_slot]);
    ++p_seq;
    if (++p_ip == C::NPL) { p_ip = 0; ++p_pi; }
    if (++p_slot == C::R) p_slot = 0;
  }
  __device__ __forceinline__ const T* wait_plane() {
    mbar_wait(&full[slot], parity);
    return ring + slot * C::PLANE_ELEMS;
  }
  __device__ __forceinline__ const T* previous_plane() const {
    return ring + (slot == 0 ? C::R - 1 : slot - 1) * C::PLANE_ELEMS;
  }
  __device__ __forceinline__ void advance_plane() {          // ring cursor only; the plane index is compile-time
    if (++slot == C::R) { slot = 0; parity ^= 1u; }
  }

  // After the group barrier of an iteration: write out the plane updated (staged) in it -- zero-based interior plane
  // `plane` of patch pi, sitting in staging buffer `buffer`.  Executed by the interior warps; thread 0 issues the TMA stores.
  __device__ __forceinline__ void drain_staged_plane(const FvGather<T>& gather, int plane, int buffer, int n_drain_threads) {
    const long long patch = g_index + (long long)pi * n_groups;
    const T* sbuf = stage + buffer * (C::STAGE_SEGS * C::SEG_PITCH);
    if (C::UNHALOED) {
      T* dst = gather.template out<C::GATHER>(q_out, patch, C::OUT_PATCH_ELEMS) + (long long)plane * C::OUT_PLANE_ELEMS;
      if (C::USE_TMA_STORE) {
        if (gt == 0) {
#pragma unroll
          for (int sgm = 0; sgm < C::STAGE_SEGS; ++sgm)
            tma_store_1d(dst + sgm * C::SEG_ELEMS, sbuf + sgm * C::SEG_PITCH, C::SEG_ELEMS * (uint32_t)sizeof(T));
          tma_store_commit();
        }
      } else {
        for (int e = gt; e < C::OUT_PLANE_ELEMS; e += n_drain_threads) {
          const int sgm = e / C::SEG_ELEMS;
          dst[e] = sbuf[sgm * C::SEG_PITCH + (e - sgm * C::SEG_ELEMS)];
        }
      }
    } else {
      // haloed layout: interior rows of the plane are runs of P*NV values (test.cpp:96-104 writes all NV)
      T* dst = gather.template out<C::GATHER>(q_out, patch, C::PATCH_ELEMS) + (long long)(plane + C::H) * C::PLANE_ELEMS;
      constexpr int ROW = C::P * C::NV;
      for (int e = gt; e < C::OUT_PLANE_ELEMS; e += n_drain_threads) {
        const int row = e / ROW;
        const int sgm = e / C::SEG_ELEMS;
        dst[((row + C::H) * C::S + C::H) * C::NV + (e - row * ROW)] = sbuf[sgm * C::SEG_PITCH + (e - sgm * C::SEG_ELEMS)];
      }
    }
  }
};

// One plane of the current patch for an interior column.  Compile-time: the phase PH = ip % 3 of the rolling window
// {old, mid, new} along axis 0 (a renaming of three register sets) and the KIND of plane -- first (halo), second (first
// interior plane: nothing to update yet), middle, last (halo: update only, publish the patch's eigenvalue).  Run-time:
// the plane index ip of middle planes and the scratch / staging buffer parity ip & 1.  Six bodies exist per kernel
// (first, second, 3 x middle, last), small enough to stay in the instruction cache with five groups at different
// places of the loop (a fully unrolled patch, 10 bodies, stalled on instruction fetch: profiles/r01_march_v3).
// Order inside one iteration:
//   load plane ip, F_0 / L_0 (registers)  ->  update plane ip-1 (reads scratch of plane ip-1, written last iteration)
//   ->  F_1, F_2, L_1, L_2 of plane ip into the other scratch buffer  ->  group barrier  ->  drain the updated plane.
// The scratch buffer written here was last read before the previous barrier, so two buffers and one barrier per plane
// are enough.  Nothing of the window survives a patch boundary (planes 0 and 1 never read `old`).
enum { MARCH_FIRST = 0, MARCH_SECOND = 1, MARCH_MIDDLE = 2, MARCH_LAST = 3 };

template <class C, int PH, int KIND>
__device__ __forceinline__ void march_interior_step(MarchStream<C>& ms, const FvGather<typename C::T>& gather, int ip,
                                                    int cell, int sj, int sk, int st,
                                                    typename C::T (&q)[3][C::NV], typename C::T (&fi)[3][C::NR],
                                                    typename C::T (&li)[3], typename C::T (&lj)[3],
                                                    typename C::T (&lk)[3], typename C::T& lam_local,
                                                    typename FloatBits<typename C::T>::type& group_lam) {
  using T = typename C::T;
  using Phys = typename C::Phys;
  using Upd = typename C::Upd;
  using Bits = typename FloatBits<T>::type;
  constexpr int NV = C::NV, NR = C::NR, SJ = C::SJ, SK = C::SK, PJ = C::PJ, S = C::S;
  constexpr int NEW = PH % 3, MID = (PH + 2) % 3, OLD = (PH + 1) % 3;
  constexpr bool INNER = (KIND == MARCH_SECOND || KIND == MARCH_MIDDLE), UPDATE = (KIND >= MARCH_MIDDLE),
                 LAST = (KIND == MARCH_LAST);
  const int wb = ip & 1, rb = wb ^ 1;        // scratch / staging buffer written / read in this iteration

  // ------------------------------------------------------------ plane ip: state, F_0, L_0 into the window
  const T* __restrict__ qs = ms.wait_plane();
#pragma unroll
  for (int v = 0; v < NV; ++v) q[NEW][v] = qs[cell * NV + v];
  const auto pr = Phys::template prims<T>(q[NEW]);
  Phys::template flux<0, T>(q[NEW], pr, fi[NEW]);
  li[NEW] = Phys::template eigen<0, T>(q[NEW], pr);

  // ------------------------------------------------------------ update plane ip-1 (needs F_0 of planes ip-2 and ip)
  if constexpr (UPDATE) {
    const T* __restrict__ qm = ms.previous_plane();      // plane ip-1: the neighbours' Q for the dissipation
    const T* __restrict__ FjR = ms.Fj + rb * (NR * SJ) + sj;
    const T* __restrict__ FkR = ms.Fk + rb * (NR * SK) + sk;
    const T* __restrict__ LjR = ms.Lj + rb * SJ + sj;
    const T* __restrict__ LkR = ms.Lk + rb * SK + sk;
    const T dt = ms.dt;
    T qc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) qc[v] = q[MID][v];
    // "Q_copy = Q_copy - 0.5*F[+1] + 0.5*F[-1]" for axis 0, 1, 2 in order (test.cpp:60-77)
#pragma unroll
    for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], fi[NEW][v], fi[OLD][v]);
#pragma unroll
    for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], FjR[v * SJ + PJ], FjR[v * SJ - PJ]);
#pragma unroll
    for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], FkR[v * SK + 1], FkR[v * SK - 1]);
    // "Q_copy = 0.5*dt*(...) + Q_copy" from the original Q, axis 0, 1, 2 in order (test.cpp:78-95)
#pragma unroll
    for (int v = 0; v < C::DV; ++v)
      qc[v] = Upd::dissipation(qc[v], q[MID][v], q[NEW][v], q[OLD][v], li[MID], li[NEW], li[OLD], dt);
    {
      const T l_plus = LjR[PJ], l_minus = LjR[-PJ];
#pragma unroll
      for (int v = 0; v < C::DV; ++v)
        qc[v] = Upd::dissipation(qc[v], q[MID][v], qm[(cell + S) * NV + v], qm[(cell - S) * NV + v], lj[MID], l_plus,
                                 l_minus, dt);
    }
    {
      const T l_plus = LkR[1], l_minus = LkR[-1];
#pragma unroll
      for (int v = 0; v < C::DV; ++v)
        qc[v] = Upd::dissipation(qc[v], q[MID][v], qm[(cell + 1) * NV + v], qm[(cell - 1) * NV + v], lk[MID], l_plus,
                                 l_minus, dt);
    }
    fv_apply_source<Phys, Upd, T>(qc, q[MID], dt);            // "Q_copy = Q_copy + dt*S" (families with a source term)
    T* dst = ms.stage + wb * (C::STAGE_SEGS * C::SEG_PITCH) + st;
#pragma unroll
    for (int v = 0; v < NV; ++v) dst[v] = qc[v];
    if (C::USE_TMA_STORE) fence_proxy_async_smem();
  }

  // ------------------------------------------------------------ plane ip: F_1, F_2, L_1, L_2 for the neighbours
  if constexpr (INNER) {
    T* __restrict__ FjW = ms.Fj + wb * (NR * SJ) + sj;
    T* __restrict__ FkW = ms.Fk + wb * (NR * SK) + sk;
    T F[NR];
    Phys::template flux<1, T>(q[NEW], pr, F);
#pragma unroll
    for (int v = 0; v < NR; ++v) FjW[v * SJ] = F[v];
    lj[NEW] = Phys::template eigen<1, T>(q[NEW], pr);
    ms.Lj[wb * SJ + sj] = lj[NEW];
    Phys::template flux<2, T>(q[NEW], pr, F);
#pragma unroll
    for (int v = 0; v < NR; ++v) FkW[v * SK] = F[v];
    lk[NEW] = Phys::template eigen<2, T>(q[NEW], pr);
    ms.Lk[wb * SK + sk] = lk[NEW];
    lam_local = fv_max(lam_local, fv_max(li[NEW], fv_max(lj[NEW], lk[NEW])));
  }
  // per-patch maximum eigenvalue over interior cells of the input state: complete at the patch's last plane
  if constexpr (LAST) {
    T m = lam_local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fv_max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((ms.gt & 31) == 0) atomicMax(&ms.lam_slot[ms.pi & 1], FloatBits<T>::to(m));
    lam_local = T(0);
  }
  if (C::USE_TMA_STORE && UPDATE && ms.gt == 0) tma_store_wait_read();   // the other staging buffer is free again
  named_barrier_sync(ms.bar_id, C::GROUP_THREADS);

  // ------------------------------------------------------------ drain, prefetch, publish
  if constexpr (UPDATE) ms.drain_staged_plane(gather, ip - 2, wb, C::FACE_BASE);
  if (ms.gt == 0) {
    // the previous plane of the stream was last read by the update above: its ring slot takes the next plane to request
    if ((KIND != MARCH_FIRST || ms.pi >= 1) && ms.p_seq < ms.n_seq) ms.issue_next_load(gather);
    if constexpr (LAST) {
      const Bits b = ms.lam_slot[ms.pi & 1];
      ms.lam_slot[ms.pi & 1] = 0;                           // next used two patches from now
      if (ms.lambda_patch) ms.lambda_patch[ms.g_index + (long long)ms.pi * ms.n_groups] = FloatBits<T>::from(b);
      group_lam = (b > group_lam) ? b : group_lam;
    }
  }
  ms.advance_plane();
}

// all planes of one patch for an interior column: first, second, middle planes 2..P (three phases in rotation), last
template <class C, class... A>
__device__ __forceinline__ void march_interior_patch(MarchStream<C>& ms, const FvGather<typename C::T>& gather, A&... a) {
  if constexpr (C::GATHER)   // CellData::dt of this patch
    if (gather.dt != nullptr) ms.dt = gather.dt[ms.g_index + (long long)ms.pi * ms.n_groups];
  march_interior_step<C, 0, MARCH_FIRST>(ms, gather, 0, a...);
  march_interior_step<C, 1, MARCH_SECOND>(ms, gather, 1, a...);
  if constexpr (C::P >= 2) {
    int ip = 2;
    while (true) {
      march_interior_step<C, 2, MARCH_MIDDLE>(ms, gather, ip, a...);
      if (++ip > C::P) break;
      march_interior_step<C, 0, MARCH_MIDDLE>(ms, gather, ip, a...);
      if (++ip > C::P) break;
      march_interior_step<C, 1, MARCH_MIDDLE>(ms, gather, ip, a...);
      if (++ip > C::P) break;
    }
  }
  march_interior_step<C, (C::P + 1) % 3, MARCH_LAST>(ms, gather, C::P + 1, a...);
}

// One plane for a face-halo column of axis 1 or 2: F_axis and L_axis of the cell one layer outside the interior.
// `axis` differs between the lanes of the face warp (f = [axis-1 low | axis-1 high | axis-2 low | axis-2 high]); both are
// evaluated by ONE instruction stream -- the flux of either axis is the same arithmetic on a different momentum
// component (physics.cuh: flux_runtime / eigen_runtime), selected per lane -- instead of two divergent halves.
template <class C>
__device__ __forceinline__ void march_face_eval(const MarchStream<C>& ms, const typename C::T* __restrict__ qs, int cell,
                                                int axis, typename C::T* __restrict__ Fw, typename C::T* __restrict__ Lw,
                                                int comp_stride) {
  using T = typename C::T;
  using Phys = typename C::Phys;
  T q[C::NV];
#pragma unroll
  for (int v = 0; v < C::NV; ++v) q[v] = qs[cell * C::NV + v];
  const auto pr = Phys::template prims<T>(q);
  T F[C::NR];
  T L;
  if constexpr (has_runtime_axis<Phys, T>::value) {
    Phys::template flux_runtime<T>(q, pr, axis, F);
    L = Phys::template eigen_runtime<T>(q, pr, axis);
  } else if (axis == 1) {     // functors without the run-time forms (generated from SymPy / user source): two divergent halves
    Phys::template flux<1, T>(q, pr, F);
    L = Phys::template eigen<1, T>(q, pr);
  } else {
    Phys::template flux<2, T>(q, pr, F);
    L = Phys::template eigen<2, T>(q, pr);
  }
#pragma unroll
  for (int v = 0; v < C::NR; ++v) Fw[v * comp_stride] = F[v];
  *Lw = L;
}

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
fv3d_march_kernel(const typename C::T* q_in, typename C::T* q_out, long long n_patches, typename C::T dt,
                  typename C::T* __restrict__ lambda_patch, typename C::T* __restrict__ lambda_max,
                  const FvGather<typename C::T> gather) {
  using T = typename C::T;
  using Bits = typename FloatBits<T>::type;
  constexpr int P = C::P, H = C::H, S = C::S, NV = C::NV, NR = C::NR, R = C::R, NPL = C::NPL;

  // device-resident time step (peer_mail.cuh): every warp derives the same dt from this device's mailbox
  dt = peer_loop_dt<T>(gather.peer, (int)(threadIdx.x & 31), dt, blockIdx.x == 0 && threadIdx.x < 32);
  extern __shared__ __align__(128) unsigned char smem[];
  const int group = threadIdx.x / C::GROUP_THREADS;
  const int gt = threadIdx.x - group * C::GROUP_THREADS;      // thread within the group
  unsigned char* const gs = smem + group * C::GROUP_BYTES;

  MarchStream<C> ms;
  ms.q_in = q_in; ms.q_out = q_out; ms.lambda_patch = lambda_patch; ms.dt = dt;
  ms.ring = reinterpret_cast<T*>(gs + C::OFF_RING);
  ms.Fj = reinterpret_cast<T*>(gs + C::OFF_FJ);          // [2][NR][SJ]
  ms.Fk = reinterpret_cast<T*>(gs + C::OFF_FK);          // [2][NR][SK]
  ms.Lj = reinterpret_cast<T*>(gs + C::OFF_LJ);          // [2][SJ]
  ms.Lk = reinterpret_cast<T*>(gs + C::OFF_LK);          // [2][SK]
  ms.stage = reinterpret_cast<T*>(gs + C::OFF_STAGE);    // [2][STAGE_SEGS * SEG_PITCH]
  ms.lam_slot = reinterpret_cast<Bits*>(gs + C::OFF_LAM);   // [2] by patch parity
  ms.full = reinterpret_cast<unsigned long long*>(gs + C::OFF_BAR);   // [R]
  ms.gt = gt;
  ms.bar_id = 1 + group;

  if (gt == 0) {
    for (int s = 0; s < R; ++s) mbar_init(&ms.full[s], 1);
    ms.lam_slot[0] = 0;
    ms.lam_slot[1] = 0;
    fence_mbar_init();
  }
  __syncthreads();   // the only CTA-wide barrier; groups are independent from here on

  ms.n_groups = (long long)gridDim.x * C::NG;
  ms.g_index = (long long)blockIdx.x * C::NG + group;
  const long long my_patches = (n_patches > ms.g_index) ? (n_patches - ms.g_index + ms.n_groups - 1) / ms.n_groups : 0;
  ms.n_seq = (int)(my_patches * NPL);
  ms.n_my_patches = (int)my_patches;
  ms.pi = ms.slot = 0;
  ms.parity = 0;
  ms.p_seq = ms.p_ip = ms.p_pi = ms.p_slot = 0;
  if (gt == 0)
    for (int s = 0; s < R && ms.p_seq < ms.n_seq; ++s) ms.issue_next_load(gather);

  if (gt < C::FACE_BASE) {
    // ======================================================== interior columns (whole warps; lanes past N_INT idle along)
    int j = 0, k = 0;
    const bool live = gt < C::N_INT;
    if (live) C::column(gt, j, k);
    const int cell = (j + H) * S + (k + H);
    const int sj = (j + 1) * C::PJ + k;
    const int sk = j * C::PK + (k + 1);
    const int st = C::stage_index(j, k);
    T q[3][NV], fi[3][NR], li[3], lj[3], lk[3];
#pragma unroll
    for (int w = 0; w < 3; ++w) {
#pragma unroll
      for (int v = 0; v < NV; ++v) q[w][v] = T(0);
#pragma unroll
      for (int v = 0; v < NR; ++v) fi[w][v] = T(0);
      li[w] = lj[w] = lk[w] = T(0);
    }
    T lam_local = T(0);
    Bits group_lam = 0;
    // lanes past N_INT (patch sizes whose P*P is not a multiple of 32) recompute column (0,0) and write the same
    // values to the same places as lane 0: harmless, and it keeps every warp whole for the shuffles and barriers
    for (; ms.pi < ms.n_my_patches; ++ms.pi)
      march_interior_patch<C>(ms, gather, cell, sj, sk, st, q, fi, li, lj, lk, lam_local, group_lam);
    if (gt == 0) {
      if (C::USE_TMA_STORE) tma_store_wait_all();
      if (lambda_max != nullptr && group_lam != 0) atomicMax(reinterpret_cast<Bits*>(lambda_max), group_lam);
    }
  } else {
    // ======================================================== face-halo columns of axes 1 and 2
    // f = 0..4P-1 -> [axis-1 low | axis-1 high | axis-2 low | axis-2 high], P columns each
    const int f = gt - C::FACE_BASE;
    const bool live = f < C::N_FACE;
    const int f_axis = (f / (2 * P)) ? 2 : 1;
    const int f_side = (f / P) & 1;
    const int f_pos = f % P;
    const int edge = f_side ? H + P : H - 1;
    const int cell = (f_axis == 1) ? edge * S + (f_pos + H) : (f_pos + H) * S + edge;
    const int slot_in_scratch = (f_axis == 1) ? (f_side ? P + 1 : 0) * C::PJ + f_pos
                                              : f_pos * C::PK + (f_side ? P + 1 : 0);
    // per lane: where this column's F / L go in the scratch of its axis (buffer 0; buffer 1 is one buffer further)
    T* const F_base = (f_axis == 1 ? ms.Fj : ms.Fk) + slot_in_scratch;
    T* const L_base = (f_axis == 1 ? ms.Lj : ms.Lk) + slot_in_scratch;
    const int comp_stride = (f_axis == 1) ? C::SJ : C::SK;
    for (; ms.pi < ms.n_my_patches; ++ms.pi) {
      for (int ip = 0; ip < NPL; ++ip) {
        if (ip >= 1 && ip <= P) {          // halo planes need no axis-1/2 fluxes: the face warps do not even wait for them
          const T* __restrict__ qs = ms.wait_plane();
          if (live) {
            const int buf = ip & 1;
            march_face_eval<C>(ms, qs, cell, f_axis, F_base + buf * (NR * comp_stride), L_base + buf * comp_stride, comp_stride);
          }
        }
        named_barrier_sync(ms.bar_id, C::GROUP_THREADS);
        ms.advance_plane();
      }
    }
  }
}

template <class C>
struct Fv3dMarchLauncher {
  static cudaError_t prepare(FvLaunchInfo* info, long long n_patches) {
    static int cached_ctas_per_sm[64];
    static int cached_sms[64];
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (cached_ctas_per_sm[dev] == 0) {
      err = cudaFuncSetAttribute(fv3d_march_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
      if (err != cudaSuccess) return err;
      int per_sm = 0, sms = 0;
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fv3d_march_kernel<C>, C::NT, C::SMEM_BYTES);
      if (err != cudaSuccess) return err;
      err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (err != cudaSuccess) return err;
      if (per_sm < 1) return cudaErrorLaunchOutOfResources;
      cached_sms[dev] = sms;
      cached_ctas_per_sm[dev] = per_sm;
    }
    const long long ctas_needed = (n_patches + C::NG - 1) / C::NG;
    const long long resident = (long long)cached_sms[dev] * cached_ctas_per_sm[dev];
    info->grid = (int)(ctas_needed < resident ? ctas_needed : resident);
    info->block = C::NT;
    info->smem_bytes = C::SMEM_BYTES;
    info->patches_per_tile = C::NG;
    info->ctas_per_sm = cached_ctas_per_sm[dev];
    return cudaSuccess;
  }

  static cudaError_t launch(const void* q_in, void* q_out, long long n_patches, double dt, void* lambda_patch,
                            void* lambda_max, cudaStream_t stream, const FvGatherRaw* gather = nullptr) {
    using T = typename C::T;
    if (n_patches <= 0) return cudaSuccess;
    FvLaunchInfo info;
    cudaError_t err = prepare(&info, n_patches);
    if (err != cudaSuccess) return err;
    fv3d_march_kernel<C><<<info.grid, info.block, info.smem_bytes, stream>>>(
        static_cast<const T*>(q_in), static_cast<T*>(q_out), n_patches, static_cast<T>(dt),
        static_cast<T*>(lambda_patch), static_cast<T*>(lambda_max), make_gather<T>(gather));
    return cudaGetLastError();
  }
};

// what a generated unit (exahype.printers.CUDAPrinter) instantiates: as many groups per CTA as 227 KB of shared memory and
// 128 registers per thread allow, planes through a 5-deep ring
template <class Phys, class Upd, typename T, int P, int H, bool DA, bool UH>
struct Fv3dMarchAutoConfig {
  using One = Fv3dMarchConfig<Phys, Upd, T, P, H, 1, 5, 1, DA, UH>;
  static constexpr int BY_SMEM = (227 * 1024) / One::GROUP_BYTES;
  static constexpr int BY_REGS = 512 / One::GROUP_THREADS;
  static constexpr int NG0 = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
  static constexpr int NG = NG0 < 1 ? 1 : (NG0 > 15 ? 15 : NG0);
  using type = Fv3dMarchConfig<Phys, Upd, T, P, H, NG, 5, 1, DA, UH>;
  using gather_type = Fv3dMarchConfig<Phys, Upd, T, P, H, NG, 5, 1, DA, UH, true>;
};
template <class Phys, class Upd, typename T, int P, int H, bool DA, bool UH>
using Fv3dMarchAuto = Fv3dMarchLauncher<typename Fv3dMarchAutoConfig<Phys, Upd, T, P, H, DA, UH>::type>;

}  // namespace exahype
