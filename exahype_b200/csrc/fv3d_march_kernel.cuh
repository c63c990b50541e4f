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
//   * only the current plane's F_1, F_2, L_1, L_2 go through shared scratch (three rotating plane buffers -> one
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

#include "fv_patch_kernel.cuh"

namespace exahype {

template <class Phys_, class Upd_, typename T_, int P_, int H_, int NG_, int R_, int MINB_, bool DISS_ALL_,
          bool UNHALOED_>
struct Fv3dMarchConfig {
  using Phys = Phys_;
  using Upd = Upd_;
  using T = T_;
  static constexpr int DIM = 3, P = P_, H = H_, NG = NG_, R = R_, MINB = MINB_;
  static constexpr bool DISS_ALL = DISS_ALL_, UNHALOED = UNHALOED_;
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
  static constexpr int OFF_FK = align_up(OFF_FJ + 3 * NR * SJ * (int)sizeof(T), 16);
  static constexpr int OFF_LJ = align_up(OFF_FK + 3 * NR * SK * (int)sizeof(T), 16);
  static constexpr int OFF_LK = align_up(OFF_LJ + 3 * SJ * (int)sizeof(T), 16);
  static constexpr int OFF_STAGE = align_up(OFF_LK + 3 * SK * (int)sizeof(T), 128);
  static constexpr int OFF_LAM = align_up(OFF_STAGE + 2 * STAGE_SEGS * SEG_PITCH * (int)sizeof(T), 16);
  static constexpr int OFF_BAR = align_up(OFF_LAM + 2 * 8, 16);
  static constexpr int GROUP_BYTES = align_up(OFF_BAR + R * 8, 128);
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

__device__ __forceinline__ void named_barrier_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
fv3d_march_kernel(const typename C::T* q_in, typename C::T* q_out, long long n_patches, typename C::T dt,
                  typename C::T* __restrict__ lambda_patch, typename C::T* __restrict__ lambda_max) {
  using T = typename C::T;
  using Phys = typename C::Phys;
  using Upd = typename C::Upd;
  using Bits = typename FloatBits<T>::type;
  constexpr int P = C::P, H = C::H, S = C::S, NV = C::NV, NR = C::NR, R = C::R, NPL = C::NPL;
  constexpr int SJ = C::SJ, SK = C::SK, PJ = C::PJ, PK = C::PK;

  extern __shared__ __align__(128) unsigned char smem[];
  const int group = threadIdx.x / C::GROUP_THREADS;
  const int gt = threadIdx.x - group * C::GROUP_THREADS;      // thread within the group
  unsigned char* const gs = smem + group * C::GROUP_BYTES;
  T* const ring = reinterpret_cast<T*>(gs + C::OFF_RING);
  T* const Fj = reinterpret_cast<T*>(gs + C::OFF_FJ);          // [3][NR][SJ]
  T* const Fk = reinterpret_cast<T*>(gs + C::OFF_FK);          // [3][NR][SK]
  T* const Lj = reinterpret_cast<T*>(gs + C::OFF_LJ);          // [3][SJ]
  T* const Lk = reinterpret_cast<T*>(gs + C::OFF_LK);          // [3][SK]
  T* const stage = reinterpret_cast<T*>(gs + C::OFF_STAGE);    // [2][STAGE_SEGS * SEG_PITCH]
  Bits* const lam_slot = reinterpret_cast<Bits*>(gs + C::OFF_LAM);   // [2] by patch parity
  unsigned long long* const full = reinterpret_cast<unsigned long long*>(gs + C::OFF_BAR);   // [R]

  if (gt == 0) {
    for (int s = 0; s < R; ++s) mbar_init(&full[s], 1);
    lam_slot[0] = 0;
    lam_slot[1] = 0;
    fence_mbar_init();
  }
  __syncthreads();   // the only CTA-wide barrier; groups are independent from here on

  const long long n_groups = (long long)gridDim.x * C::NG;
  const long long g_index = (long long)blockIdx.x * C::NG + group;
  const long long my_patches = (n_patches > g_index) ? (n_patches - g_index + n_groups - 1) / n_groups : 0;
  const long long n_seq = my_patches * NPL;                   // planes this group streams
  const int bar_id = 1 + group;

  // plane `seq` of this group's stream -> global source
  auto issue_load = [&](long long seq) {
    const long long pi = seq / NPL;
    const int ip = (int)(seq - pi * NPL);
    const long long patch = g_index + pi * n_groups;
    const int slot = (int)(seq % R);
    mbar_expect_tx(&full[slot], C::PLANE_BYTES);
    tma_load_1d(ring + slot * C::PLANE_ELEMS,
                q_in + patch * (long long)C::PATCH_ELEMS + (long long)(ip + H - 1) * C::PLANE_ELEMS,
                C::PLANE_BYTES, &full[slot]);
  };
  if (gt == 0)
    for (long long s = 0; s < R && s < n_seq; ++s) issue_load(s);

  // ---- roles
  const bool is_interior = gt < C::N_INT;
  const bool is_face = gt >= C::FACE_BASE && gt < C::FACE_BASE + C::N_FACE;
  int j = 0, k = 0;
  if (is_interior) C::column(gt, j, k);
  // face columns: f = 0..4P-1 -> [axis-1 low | axis-1 high | axis-2 low | axis-2 high], P columns each
  const int f = gt - C::FACE_BASE;
  const int f_axis = (f / (2 * P)) ? 2 : 1;
  const int f_side = (f / P) & 1;
  const int f_pos = f % P;
  // cell within a haloed plane and scratch slot of this thread's column
  const int cell = is_interior ? (j + H) * S + (k + H)
                 : (f_axis == 1 ? (f_side ? H + P : H - 1) * S + (f_pos + H)
                                : (f_pos + H) * S + (f_side ? H + P : H - 1));
  const int sj = is_interior ? (j + 1) * PJ + k : (f_side ? P + 1 : 0) * PJ + f_pos;      // axis-1 scratch slot
  const int sk = is_interior ? j * PK + (k + 1) : f_pos * PK + (f_side ? P + 1 : 0);      // axis-2 scratch slot
  const int st = is_interior ? C::stage_index(j, k) : 0;

  // ---- rolling window along axis 0 (interior columns)
  T q_old[NV], q_mid[NV], q_new[NV];        // cell state of planes ip-2, ip-1, ip
  T fi_old[NR], fi_mid[NR], fi_new[NR];     // F_0 of the same planes
  T li_old = T(0), li_mid = T(0), li_new = T(0);
  T lj_mid = T(0), lk_mid = T(0), lj_new = T(0), lk_new = T(0);
#pragma unroll
  for (int v = 0; v < NV; ++v) q_old[v] = q_mid[v] = q_new[v] = T(0);
#pragma unroll
  for (int v = 0; v < NR; ++v) fi_old[v] = fi_mid[v] = fi_new[v] = T(0);
  T lam_local = T(0);
  Bits group_lam = 0;

  long long pi = 0;     // patch counter of this group
  int ip = 0;           // plane within the patch, 0 .. P+1
  for (long long seq = 0; seq <= n_seq; ++seq) {
    const int buf = (int)(seq % 3);
    const bool have_plane = seq < n_seq;      // the extra iteration only drains the last staged plane
    if (have_plane) {
      const int slot = (int)(seq % R);
      mbar_wait(&full[slot], (uint32_t)((seq / R) & 1));
      const T* __restrict__ qs = ring + slot * C::PLANE_ELEMS;
      const bool inner_plane = (ip >= 1 && ip <= P);

      // ------------------------------------------------------------ evaluate plane ip
      if (is_interior) {
#pragma unroll
        for (int v = 0; v < NV; ++v) q_new[v] = qs[cell * NV + v];
        const auto pr = Phys::template prims<T>(q_new);
        Phys::template flux<0, T>(q_new, pr, fi_new);
        li_new = Phys::template eigen<0, T>(q_new, pr);
        if (inner_plane) {
          T F[NR];
          Phys::template flux<1, T>(q_new, pr, F);
#pragma unroll
          for (int v = 0; v < NR; ++v) Fj[(buf * NR + v) * SJ + sj] = F[v];
          lj_new = Phys::template eigen<1, T>(q_new, pr);
          Lj[buf * SJ + sj] = lj_new;
          Phys::template flux<2, T>(q_new, pr, F);
#pragma unroll
          for (int v = 0; v < NR; ++v) Fk[(buf * NR + v) * SK + sk] = F[v];
          lk_new = Phys::template eigen<2, T>(q_new, pr);
          Lk[buf * SK + sk] = lk_new;
          lam_local = fv_max(lam_local, fv_max(li_new, fv_max(lj_new, lk_new)));
        }
      } else if (is_face && inner_plane) {
        T q[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) q[v] = qs[cell * NV + v];
        const auto pr = Phys::template prims<T>(q);
        T F[NR];
        if (f_axis == 1) {
          Phys::template flux<1, T>(q, pr, F);
#pragma unroll
          for (int v = 0; v < NR; ++v) Fj[(buf * NR + v) * SJ + sj] = F[v];
          Lj[buf * SJ + sj] = Phys::template eigen<1, T>(q, pr);
        } else {
          Phys::template flux<2, T>(q, pr, F);
#pragma unroll
          for (int v = 0; v < NR; ++v) Fk[(buf * NR + v) * SK + sk] = F[v];
          Lk[buf * SK + sk] = Phys::template eigen<2, T>(q, pr);
        }
      }
      // per-patch maximum eigenvalue over interior cells of the input state: published at the patch's last plane
      if (ip == NPL - 1 && is_interior) {
        T m = lam_local;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fv_max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((gt & 31) == 0) atomicMax(&lam_slot[pi & 1], FloatBits<T>::to(m));
        lam_local = T(0);
      }
    }
    if (C::USE_TMA_STORE && gt == 0) tma_store_wait_read();   // staging buffer (seq & 1) is free again
    named_barrier_sync(bar_id, C::GROUP_THREADS);

    // ------------------------------------------------------------ after the barrier: drain + prefetch (one thread)
    // the plane staged in the previous iteration belongs to (patch, plane) = previous (pi, ip) - 1
    const bool staged_prev = (seq >= 1) && ((ip == 0) ? true : (ip >= 3));   // previous iteration had ip_prev >= 2
    long long prev_pi = pi;
    int prev_plane = ip - 3;            // zero-based interior plane written in the previous iteration
    if (ip == 0) { prev_pi = pi - 1; prev_plane = P - 1; }
    if (staged_prev && prev_pi >= 0) {
      const long long patch = g_index + prev_pi * n_groups;
      const T* sbuf = stage + ((seq - 1) & 1) * (C::STAGE_SEGS * C::SEG_PITCH);
      if (C::UNHALOED) {
        T* dst = q_out + patch * (long long)C::OUT_PATCH_ELEMS + (long long)prev_plane * C::OUT_PLANE_ELEMS;
        if (C::USE_TMA_STORE) {
          if (gt == 0) {
#pragma unroll
            for (int sgm = 0; sgm < C::STAGE_SEGS; ++sgm)
              tma_store_1d(dst + sgm * C::SEG_ELEMS, sbuf + sgm * C::SEG_PITCH, C::SEG_ELEMS * (uint32_t)sizeof(T));
            tma_store_commit();
          }
        } else {
          for (int e = gt; e < C::OUT_PLANE_ELEMS; e += C::GROUP_THREADS) {
            const int sgm = e / C::SEG_ELEMS;
            dst[e] = sbuf[sgm * C::SEG_PITCH + (e - sgm * C::SEG_ELEMS)];
          }
        }
      } else {
        // haloed layout: interior rows of the plane are runs of P*NV values (test.cpp:96-104 writes all NV)
        T* dst = q_out + patch * (long long)C::PATCH_ELEMS + (long long)(prev_plane + H) * C::PLANE_ELEMS;
        constexpr int ROW = P * NV;
        for (int e = gt; e < C::OUT_PLANE_ELEMS; e += C::GROUP_THREADS) {
          const int row = e / ROW;
          const int sgm = e / C::SEG_ELEMS;
          dst[((row + H) * S + H) * NV + (e - row * ROW)] = sbuf[sgm * C::SEG_PITCH + (e - sgm * C::SEG_ELEMS)];
        }
      }
    }
    if (gt == 0) {
      if (seq >= 2 && seq - 2 + R < n_seq) issue_load(seq - 2 + R);   // slot of plane seq-2 was last read in iteration seq-1
      if (have_plane && ip == 0 && pi >= 1) {                          // previous patch is complete: publish its lambda
        const Bits b = lam_slot[(pi - 1) & 1];
        lam_slot[(pi - 1) & 1] = 0;
        if (lambda_patch) lambda_patch[g_index + (pi - 1) * n_groups] = FloatBits<T>::from(b);
        group_lam = (b > group_lam) ? b : group_lam;
      }
    }
    if (!have_plane) break;

    // ------------------------------------------------------------ update plane ip-1 (needs F_0 of planes ip-2 and ip)
    if (is_interior && ip >= 2) {
      const int ub = (int)((seq + 2) % 3);      // scratch buffer of plane ip-1, written in the previous iteration
      const T* __restrict__ qm = ring + (int)((seq - 1) % R) * C::PLANE_ELEMS;   // plane ip-1, for the neighbours' Q
      T qc[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) qc[v] = q_mid[v];
      // "Q_copy = Q_copy - 0.5*F[+1] + 0.5*F[-1]" for axis 0, 1, 2 in order (test.cpp:60-77)
#pragma unroll
      for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], fi_new[v], fi_old[v]);
#pragma unroll
      for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], Fj[(ub * NR + v) * SJ + sj + PJ], Fj[(ub * NR + v) * SJ + sj - PJ]);
#pragma unroll
      for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], Fk[(ub * NR + v) * SK + sk + 1], Fk[(ub * NR + v) * SK + sk - 1]);
      // "Q_copy = 0.5*dt*(...) + Q_copy" from the original Q, axis 0, 1, 2 in order (test.cpp:78-95)
#pragma unroll
      for (int v = 0; v < C::DV; ++v)
        qc[v] = Upd::dissipation(qc[v], q_mid[v], q_new[v], q_old[v], li_mid, li_new, li_old, dt);
      {
        const T l_plus = Lj[ub * SJ + sj + PJ], l_minus = Lj[ub * SJ + sj - PJ];
#pragma unroll
        for (int v = 0; v < C::DV; ++v)
          qc[v] = Upd::dissipation(qc[v], q_mid[v], qm[(cell + S) * NV + v], qm[(cell - S) * NV + v], lj_mid, l_plus,
                                   l_minus, dt);
      }
      {
        const T l_plus = Lk[ub * SK + sk + 1], l_minus = Lk[ub * SK + sk - 1];
#pragma unroll
        for (int v = 0; v < C::DV; ++v)
          qc[v] = Upd::dissipation(qc[v], q_mid[v], qm[(cell + 1) * NV + v], qm[(cell - 1) * NV + v], lk_mid, l_plus,
                                   l_minus, dt);
      }
      T* dst = stage + (seq & 1) * (C::STAGE_SEGS * C::SEG_PITCH) + st;
#pragma unroll
      for (int v = 0; v < NV; ++v) dst[v] = qc[v];
      if (C::USE_TMA_STORE) fence_proxy_async_smem();
    }
    // ------------------------------------------------------------ rotate the window
    if (is_interior) {
#pragma unroll
      for (int v = 0; v < NV; ++v) { q_old[v] = q_mid[v]; q_mid[v] = q_new[v]; }
#pragma unroll
      for (int v = 0; v < NR; ++v) { fi_old[v] = fi_mid[v]; fi_mid[v] = fi_new[v]; }
      li_old = li_mid; li_mid = li_new;
      lj_mid = lj_new; lk_mid = lk_new;
    }
    if (++ip == NPL) { ip = 0; ++pi; }
  }

  // the extra iteration above staged nothing new; publish the last patch's lambda and the group maximum
  if (gt == 0) {
    if (C::USE_TMA_STORE) tma_store_wait_all();
    if (my_patches > 0) {
      const Bits b = lam_slot[(my_patches - 1) & 1];
      if (lambda_patch) lambda_patch[g_index + (my_patches - 1) * n_groups] = FloatBits<T>::from(b);
      group_lam = (b > group_lam) ? b : group_lam;
    }
    if (lambda_max != nullptr && group_lam != 0) atomicMax(reinterpret_cast<Bits*>(lambda_max), group_lam);
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
                            void* lambda_max, cudaStream_t stream) {
    using T = typename C::T;
    if (n_patches <= 0) return cudaSuccess;
    FvLaunchInfo info;
    cudaError_t err = prepare(&info, n_patches);
    if (err != cudaSuccess) return err;
    fv3d_march_kernel<C><<<info.grid, info.block, info.smem_bytes, stream>>>(
        static_cast<const T*>(q_in), static_cast<T*>(q_out), n_patches, static_cast<T>(dt),
        static_cast<T*>(lambda_patch), static_cast<T*>(lambda_max));
    return cudaGetLastError();
  }
};

}  // namespace exahype
