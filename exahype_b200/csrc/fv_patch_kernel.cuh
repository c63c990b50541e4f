// sm_100a kernel template for ExaHyPE's batched stateless finite-volume Rusanov patch update.
//
// What it computes -- per patch, exactly the statement sequence the reference's CPPPrinter emits for
// examples/Batched_stateless.py (reference "Unit test/test.cpp":11-104, loop ranges per
// exahype/printers/CPPPrinter.py:116-137):
//     Qc = Q;  F_n = Flux(Qc, n), L_n = maxEigenvalue(Qc, n) on {full along n, interior across};
//     for n: Qc -= 0.5 F_n[c+e_n] - 0.5 F_n[c-e_n];   for n: Qc += 0.5 dt (...max(L_n)...) from the ORIGINAL Q;
//     Q[interior] = Qc[interior]          (+ per-patch / global max eigenvalue of the input state)
// -- but as ONE fused pass: the reference makes ten sweeps over five heap temporaries per call
// (test.cpp:4-8), here nothing but Q is read from and nothing but the interior is written to HBM.
//
// Execution model (one persistent CTA per resident slot, grid = SMs x CTAs/SM, tiles strided over the grid):
//   load   a tile = G consecutive patches is one contiguous block of the AoS batch; a single elected thread
//          brings it into shared memory with a 1-D TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx),
//          double buffered so tile t+1 streams in from HBM while tile t is computed;
//   phase A  thread <-> interior cell: variables to registers, per-cell primitives once, F_n / L_n for every
//          axis into per-axis shared scratch (SoA, compact "full along n, interior across" boxes);
//          the 2*dim*P^(dim-1) face-halo cells are evaluated for their one axis by the first warps
//          (axis is warp-uniform for the committed shapes);  per-patch max(L) by warp shuffle + smem atomic;
//   phase B  thread <-> same interior cell: neighbours' F_n, L_n (scratch) and Q (staged input) -> updated
//          cell in registers -> AoS output staging in shared memory;
//   phase C  staging -> HBM: coalesced row segments for the haloed in-place form, or one TMA bulk store
//          for the un-haloed contiguous form.
// Two CTA barriers per tile.  No global temporaries, no tensor cores (bandwidth-bound stencil).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "peer_mail.cuh"
#include "physics.cuh"

namespace exahype {

__host__ __device__ constexpr int ipow(int b, int e) { return e <= 0 ? 1 : b * ipow(b, e - 1); }
__host__ __device__ constexpr int align_up(int x, int a) { return (x + a - 1) / a * a; }

template <typename T> struct FloatBits;
template <> struct FloatBits<double> {
  using type = unsigned long long;
  static __device__ __forceinline__ type to(double x) { return (type)__double_as_longlong(x); }
  static __device__ __forceinline__ double from(type b) { return __longlong_as_double((long long)b); }
};
template <> struct FloatBits<float> {
  using type = unsigned int;
  static __device__ __forceinline__ type to(float x) { return __float_as_uint(x); }
  static __device__ __forceinline__ float from(type b) { return __uint_as_float(b); }
};

// ExaHyPE2 CellData form of a batch (reference examples/kernel-generator.py:8-19, the QIn / QOut / dt members of
// `::exahype2::CellData&`): patch p lives at q_in[p] / q_out[p] instead of base + p * stride and advances by its own
// dt[p] (null dt: the scalar).  Kernels are instantiated twice, GATHER = false (dense batch: the struct is dead and
// costs nothing; a run-time switch measured 3.5 % on the dense path) and GATHER = true.
template <typename T>
struct FvGather {
  const T* const* q_in = nullptr;
  T* const* q_out = nullptr;
  const T* dt = nullptr;
  // CellData::cellCentre / cellSize (dim values per patch) and CellData::t, for functors with the ExaHyPE2 solver
  // signature flux(Q, x, h, t, dt, normal, F) (FvCellCtx below); null: centre 0, size 1, t 0
  const T* cell_centre = nullptr;
  const T* cell_size = nullptr;
  const T* t = nullptr;
  // the step's time-step source and the all-reduce(max) of lambda_max (peer_mail.cuh): every kernel honours
  // peer.dt_in (device-resident dt); kernels whose launch info says fused_allreduce also run the exchange themselves
  // (peer.mode).  Rides along here because this struct already reaches every kernel.
  FvPeerFuse peer;
  template <bool GATHER>
  __device__ __forceinline__ const T* in(const T* base, long long patch, int patch_elems) const {
    if constexpr (GATHER) return q_in[patch];
    else return base + patch * (long long)patch_elems;
  }
  template <bool GATHER>
  __device__ __forceinline__ T* out(T* base, long long patch, int patch_elems) const {
    if constexpr (GATHER) return q_out[patch];
    else return base + patch * (long long)patch_elems;
  }
  template <bool GATHER>
  __device__ __forceinline__ T step(T scalar, long long patch) const {
    if constexpr (GATHER) return dt ? dt[patch] : scalar;
    else return scalar;
  }
};
struct FvGatherRaw {   // type-erased form crossing the registry's function pointers
  const void* const* q_in;
  void* const* q_out;
  const void* dt;
  FvPeerFuse peer;
  const void* cell_centre = nullptr;
  const void* cell_size = nullptr;
  const void* t = nullptr;
};
template <typename T>
inline FvGather<T> make_gather(const FvGatherRaw* raw) {
  FvGather<T> g;
  if (raw) {
    g.q_in = reinterpret_cast<const T* const*>(raw->q_in);
    g.q_out = reinterpret_cast<T* const*>(raw->q_out);
    g.dt = static_cast<const T*>(raw->dt);
    g.peer = raw->peer;
    g.cell_centre = static_cast<const T*>(raw->cell_centre);
    g.cell_size = static_cast<const T*>(raw->cell_size);
    g.t = static_cast<const T*>(raw->t);
  }
  return g;
}

// ------------------------------------------------------------------------------------------------
// compile-time geometry of one instantiation
template <class Phys_, class Upd_, typename T_, int DIM_, int P_, int H_, int G_, int NT_, int MINB_,
          bool DISS_ALL_, bool UNHALOED_, bool GATHER_ = false>
struct FvKernelConfig {
  using Phys = Phys_;
  using Upd = Upd_;
  using T = T_;
  static constexpr int DIM = DIM_, P = P_, H = H_, G = G_, NT = NT_, MINB = MINB_;
  static constexpr bool DISS_ALL = DISS_ALL_, UNHALOED = UNHALOED_, GATHER = GATHER_;
  static_assert(DIM == 2 || DIM == 3, "dim");
  static_assert(P >= 1 && H >= 1 && G >= 1, "patch geometry");
  static_assert(NT % 32 == 0 && NT <= 1024, "threads per CTA");

  static constexpr int NR = Phys::NR, NA = Phys::NA, NV = NR + NA;
  static constexpr int S = P + 2 * H;                 // haloed side
  static constexpr int NCELL = ipow(S, DIM);          // haloed cells per patch
  static constexpr int PD = ipow(P, DIM);             // interior cells per patch
  static constexpr int PF = ipow(P, DIM - 1);         // cells per face
  static constexpr int PATCH_ELEMS = NCELL * NV;
  static constexpr int PATCH_BYTES = PATCH_ELEMS * (int)sizeof(T);
  static constexpr int TILE_ELEMS = G * PATCH_ELEMS;
  static constexpr int SLOTS = (P + 2) * PF;          // scratch cells per axis per patch
  static constexpr int INT_CELLS = G * PD;            // interior cells per tile
  static constexpr int CPT = (INT_CELLS + NT - 1) / NT;   // interior cells per thread
  static constexpr int FACES_PER_AXIS = G * 2 * PF;
  static constexpr int FACE_CELLS = DIM * FACES_PER_AXIS;
  static constexpr int DV = DISS_ALL ? NR : 1;        // variables that receive dissipation
  static constexpr int OUT_PATCH_ELEMS = PD * NV;     // staged output per patch (aux copied through)
  static constexpr int OUT_ELEMS = G * OUT_PATCH_ELEMS;

  // TMA bulk copies need 16-byte granularity; otherwise a cooperative copy is used
  static constexpr bool USE_TMA_LOAD = (PATCH_BYTES % 16 == 0);
  static constexpr bool USE_TMA_STORE = UNHALOED && ((OUT_PATCH_ELEMS * (int)sizeof(T)) % 16 == 0);
  static constexpr int Q_BUFFERS = USE_TMA_LOAD ? 2 : 1;
  // An even AoS cell stride puts the lanes of a warp on few banks; then the neighbour values the dissipation
  // needs are stashed SoA in phase A instead of being read from the staged AoS block.
  static constexpr bool STASH_Q = (NV % 2 == 0);

  static constexpr int OFF_Q = 0;
  static constexpr int OFF_F = align_up(OFF_Q + Q_BUFFERS * TILE_ELEMS * (int)sizeof(T), 16);
  static constexpr int OFF_L = align_up(OFF_F + DIM * NR * G * SLOTS * (int)sizeof(T), 16);
  static constexpr int OFF_R = align_up(OFF_L + DIM * G * SLOTS * (int)sizeof(T), 16);
  static constexpr int OFF_OUT = align_up(OFF_R + (STASH_Q ? DV * G * NCELL * (int)sizeof(T) : 0), 128);
  static constexpr int OFF_LAM = align_up(OFF_OUT + OUT_ELEMS * (int)sizeof(T), 16);
  static constexpr int OFF_BAR = align_up(OFF_LAM + 2 * G * 8, 16);
  static constexpr int SMEM_BYTES = OFF_BAR + 2 * 8;
  static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit the 227 KB of shared memory per CTA");

  // haloed cell stride of axis m (axis 0 slowest; reference CPPPrinter.py:247-261)
  static __host__ __device__ constexpr int cell_stride(int m) { return ipow(S, DIM - 1 - m); }
  // scratch box of axis n: extent P+2 along n, P across; row-major, last axis fastest
  static __host__ __device__ constexpr int box_extent(int n, int m) { return m == n ? P + 2 : P; }
  static __host__ __device__ constexpr int box_stride(int n, int m) {
    int s = 1;
    for (int k = m + 1; k < DIM; ++k) s *= box_extent(n, k);
    return s;
  }
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk tensor-memory-accelerator copies (SASS: UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Where and when a cell is: what ExaHyPE2's solver functions take next to Q -- flux(Q, x, h, t, dt, normal, F),
// maxEigenvalue(Q, x, h, t, dt, normal) (reference "Unit test/correctness_test.cpp":90-99; call sites declared by
// examples/kernel-generator.py:37-39).  Functor families that set `static constexpr bool NEEDS_CONTEXT = true` receive it
// as an extra argument of flux<N> / eigen<N>; the committed Euler / shallow-water families do not depend on position or
// time and never see it.  Definitions, in exactly this evaluation order (the C++ side of the parity test uses the same):
//   h[d] = H[d] / patch_size                                   getVolumeSize(cellSize, patch_size)
//   x[d] = (X[d] - 0.5 * H[d]) + (index[d] + 0.5) * h[d]       getVolumeCentre(cellCentre, cellSize, patch_size, {i, j(, k)})
// with index[d] the haloed loop index the declaration passes (`{i, j}`), X / H the patch's cellCentre / cellSize.
template <typename T, int DIM>
struct FvCellCtx {
  T x[DIM], h[DIM];   // volume centre and size
  T X[DIM], H[DIM];   // patch centre and size
  T t, dt;
};
template <class Phys, class = void> struct needs_context : std::false_type {};
template <class Phys> struct needs_context<Phys, std::enable_if_t<Phys::NEEDS_CONTEXT>> : std::true_type {};

template <class C>
__device__ __forceinline__ FvCellCtx<typename C::T, C::DIM> cell_context(const FvGather<typename C::T>& gather,
                                                                        long long patch, int cell, typename C::T dt) {
  using T = typename C::T;
  FvCellCtx<T, C::DIM> ctx;
#pragma unroll
  for (int d = 0; d < C::DIM; ++d) {
    const int index = (cell / C::cell_stride(d)) % C::S;
    ctx.X[d] = gather.cell_centre ? gather.cell_centre[patch * C::DIM + d] : T(0);
    ctx.H[d] = gather.cell_size ? gather.cell_size[patch * C::DIM + d] : T(1);
    ctx.h[d] = ctx.H[d] / T(C::P);
    ctx.x[d] = (ctx.X[d] - T(0.5) * ctx.H[d]) + (T(index) + T(0.5)) * ctx.h[d];
  }
  ctx.t = gather.t ? gather.t[patch] : T(0);
  ctx.dt = dt;
  return ctx;
}
template <class C, int N, class Pr, class Ctx>
__device__ __forceinline__ void ctx_flux(const typename C::T (&q)[C::NV], const Pr& pr, typename C::T (&F)[C::NR],
                                         const Ctx& ctx) {
  if constexpr (needs_context<typename C::Phys>::value) C::Phys::template flux<N, typename C::T>(q, pr, F, ctx);
  else C::Phys::template flux<N, typename C::T>(q, pr, F);
}
template <class C, int N, class Pr, class Ctx>
__device__ __forceinline__ typename C::T ctx_eigen(const typename C::T (&q)[C::NV], const Pr& pr, const Ctx& ctx) {
  if constexpr (needs_context<typename C::Phys>::value) return C::Phys::template eigen<N, typename C::T>(q, pr, ctx);
  else return C::Phys::template eigen<N, typename C::T>(q, pr);
}

template <class C>
struct CellIndex {
  int g;             // patch within the tile
  int cell;          // haloed cell index within the patch
  int slot[C::DIM];  // scratch index within the patch, per axis
};

// interior cell e in [0, G*PD)
template <class C>
__device__ __forceinline__ CellIndex<C> interior_cell(int e) {
  CellIndex<C> ci;
  ci.g = e / C::PD;
  int r = e - ci.g * C::PD;
  ci.cell = 0;
#pragma unroll
  for (int n = 0; n < C::DIM; ++n) ci.slot[n] = 0;
#pragma unroll
  for (int m = C::DIM - 1; m >= 0; --m) {
    const int x = r % C::P;
    r /= C::P;
    ci.cell += (x + C::H) * C::cell_stride(m);
#pragma unroll
    for (int n = 0; n < C::DIM; ++n) ci.slot[n] += (x + (m == n ? 1 : 0)) * C::box_stride(n, m);
  }
  return ci;
}

// face-halo cell f in [0, FACES_PER_AXIS) of axis N: one layer outside the interior on either side
template <class C, int N>
__device__ __forceinline__ void face_cell(int f, int& g, int& cell, int& slot) {
  g = f / (2 * C::PF);
  int r = f - g * (2 * C::PF);
  const int side = r / C::PF;
  r -= side * C::PF;
  cell = (side ? C::H + C::P : C::H - 1) * C::cell_stride(N);
  slot = (side ? C::P + 1 : 0) * C::box_stride(N, N);
#pragma unroll
  for (int m = C::DIM - 1; m >= 0; --m) {
    if (m == N) continue;
    const int x = r % C::P;
    r /= C::P;
    cell += (x + C::H) * C::cell_stride(m);
    slot += x * C::box_stride(N, m);
  }
}

template <class C, int N>
__device__ __forceinline__ void eval_face(int f, int npatch, const typename C::T* __restrict__ qs,
                                          typename C::T* __restrict__ Fs, typename C::T* __restrict__ Ls,
                                          typename C::T* __restrict__ Rs, const FvGather<typename C::T>& gather,
                                          long long first_patch, typename C::T dt) {
  using T = typename C::T;
  using Phys = typename C::Phys;
  int g, cell, slot;
  face_cell<C, N>(f, g, cell, slot);
  if (g >= npatch) return;
  T q[C::NV];
  const T* src = qs + (g * C::NCELL + cell) * C::NV;
#pragma unroll
  for (int v = 0; v < C::NV; ++v) q[v] = src[v];
  const auto pr = Phys::template prims<T>(q);
  FvCellCtx<T, C::DIM> ctx;
  if constexpr (needs_context<Phys>::value)
    ctx = cell_context<C>(gather, first_patch + g, cell, gather.template step<C::GATHER>(dt, first_patch + g));
  T F[C::NR];
  ctx_flux<C, N>(q, pr, F, ctx);
  const int s = g * C::SLOTS + slot;
#pragma unroll
  for (int v = 0; v < C::NR; ++v) Fs[(N * C::NR + v) * (C::G * C::SLOTS) + s] = F[v];
  Ls[N * (C::G * C::SLOTS) + s] = ctx_eigen<C, N>(q, pr, ctx);
  if (C::STASH_Q) {
#pragma unroll
    for (int v = 0; v < C::DV; ++v) Rs[v * (C::G * C::NCELL) + g * C::NCELL + cell] = q[v];
  }
}

// ------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
fv_step_kernel(const typename C::T* q_in, typename C::T* q_out, long long n_patches,   // q_out may alias q_in
               typename C::T dt, typename C::T* __restrict__ lambda_patch, typename C::T* __restrict__ lambda_max,
               const FvGather<typename C::T> gather) {
  using T = typename C::T;
  using Phys = typename C::Phys;
  using Upd = typename C::Upd;
  using Bits = typename FloatBits<T>::type;
  constexpr int DIM = C::DIM, NV = C::NV, NR = C::NR, G = C::G, NT = C::NT, CPT = C::CPT;
  constexpr int FSTRIDE = G * C::SLOTS;   // elements per (axis, variable) plane of the flux scratch

  // device-resident time step (peer_mail.cuh): every warp derives the same dt from this device's mailbox
  dt = peer_loop_dt<T>(gather.peer, (int)(threadIdx.x & 31), dt, blockIdx.x == 0 && threadIdx.x < 32);
  extern __shared__ __align__(128) unsigned char smem[];
  T* const qbuf = reinterpret_cast<T*>(smem + C::OFF_Q);
  T* const Fs = reinterpret_cast<T*>(smem + C::OFF_F);
  T* const Ls = reinterpret_cast<T*>(smem + C::OFF_L);
  T* const Rs = reinterpret_cast<T*>(smem + C::OFF_R);
  T* const stage = reinterpret_cast<T*>(smem + C::OFF_OUT);
  Bits* const lam_bits = reinterpret_cast<Bits*>(smem + C::OFF_LAM);   // [2][G], by tile parity
  unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(smem + C::OFF_BAR);

  const int tid = threadIdx.x;
  const long long n_tiles = (n_patches + G - 1) / G;

  if (tid == 0 && C::USE_TMA_LOAD) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }
  if (tid < 2 * G) lam_bits[tid] = 0;
  __syncthreads();

  // one bulk copy per tile of a dense batch; one per patch when the patches are gathered through pointers
  auto request_tile = [&](long long t, int b) {
    const long long np = (n_patches - t * G < G) ? (n_patches - t * G) : G;
    const uint32_t bytes = (uint32_t)np * C::PATCH_BYTES;
    mbar_expect_tx(&mbar[b], bytes);
    if constexpr (!C::GATHER) {
      tma_load_1d(qbuf + b * C::TILE_ELEMS, q_in + t * (long long)C::TILE_ELEMS, bytes, &mbar[b]);
    } else {
      for (int g = 0; g < (int)np; ++g)
        tma_load_1d(qbuf + b * C::TILE_ELEMS + g * C::PATCH_ELEMS, gather.q_in[t * G + g], C::PATCH_BYTES, &mbar[b]);
    }
  };

  long long tile = blockIdx.x;
  if (C::USE_TMA_LOAD && tid == 0 && tile < n_tiles) request_tile(tile, 0);

  Bits cta_lam = 0;   // running maximum of the patches this thread published (threads < G)

  for (int it = 0; tile < n_tiles; ++it, tile += gridDim.x) {
    const int npatch = (int)((n_patches - tile * G < G) ? (n_patches - tile * G) : G);
    const int buf = C::USE_TMA_LOAD ? (it & 1) : 0;
    const int par = it & 1;
    const T* __restrict__ qs = qbuf + buf * C::TILE_ELEMS;

    if (C::USE_TMA_LOAD) {
      if (tid == 0) {
        const long long next = tile + gridDim.x;
        if (next < n_tiles) request_tile(next, buf ^ 1);   // buffer buf^1 was last read in phase B of the previous tile
      }
      mbar_wait(&mbar[buf], (uint32_t)((it >> 1) & 1));
    } else {
      for (int i = tid; i < npatch * C::PATCH_ELEMS; i += NT) {
        const int g = i / C::PATCH_ELEMS;
        qbuf[i] = gather.template in<C::GATHER>(q_in, tile * G + g, C::PATCH_ELEMS)[i - g * C::PATCH_ELEMS];
      }
      __syncthreads();
    }

    // ---------------------------------------------------------------- phase A: F_n, L_n of every needed cell
    T q[CPT][NV];
    T lam[CPT][DIM];
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
      const int e = tid + r * NT;
      const bool valid = (e < npatch * C::PD);
      T lmax = T(0);
      int g_of_cell = 0;
      if (valid) {
        const CellIndex<C> ci = interior_cell<C>(e);
        g_of_cell = ci.g;
        const T* src = qs + (ci.g * C::NCELL + ci.cell) * NV;
#pragma unroll
        for (int v = 0; v < NV; ++v) q[r][v] = src[v];
        const auto pr = Phys::template prims<T>(q[r]);
        FvCellCtx<T, DIM> ctx;
        if constexpr (needs_context<Phys>::value)
          ctx = cell_context<C>(gather, tile * G + ci.g, ci.cell, gather.template step<C::GATHER>(dt, tile * G + ci.g));
        T F[NR];
        {
          ctx_flux<C, 0>(q[r], pr, F, ctx);
          const int s = ci.g * C::SLOTS + ci.slot[0];
#pragma unroll
          for (int v = 0; v < NR; ++v) Fs[(0 * NR + v) * FSTRIDE + s] = F[v];
          lam[r][0] = ctx_eigen<C, 0>(q[r], pr, ctx);
          Ls[0 * FSTRIDE + s] = lam[r][0];
        }
        {
          ctx_flux<C, 1>(q[r], pr, F, ctx);
          const int s = ci.g * C::SLOTS + ci.slot[1];
#pragma unroll
          for (int v = 0; v < NR; ++v) Fs[(1 * NR + v) * FSTRIDE + s] = F[v];
          lam[r][1] = ctx_eigen<C, 1>(q[r], pr, ctx);
          Ls[1 * FSTRIDE + s] = lam[r][1];
        }
        if constexpr (DIM == 3) {
          ctx_flux<C, 2>(q[r], pr, F, ctx);
          const int s = ci.g * C::SLOTS + ci.slot[2];
#pragma unroll
          for (int v = 0; v < NR; ++v) Fs[(2 * NR + v) * FSTRIDE + s] = F[v];
          lam[r][2] = ctx_eigen<C, 2>(q[r], pr, ctx);
          Ls[2 * FSTRIDE + s] = lam[r][2];
        }
        if (C::STASH_Q) {
#pragma unroll
          for (int v = 0; v < C::DV; ++v) Rs[v * (G * C::NCELL) + ci.g * C::NCELL + ci.cell] = q[r][v];
        }
#pragma unroll
        for (int n = 0; n < DIM; ++n) lmax = fv_max(lmax, lam[r][n]);
      }
      // per-patch maximum eigenvalue of the input state over interior cells
      if (G == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fv_max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if ((tid & 31) == 0) atomicMax(&lam_bits[par * G], FloatBits<T>::to(lmax));
      } else if (valid) {
        atomicMax(&lam_bits[par * G + g_of_cell], FloatBits<T>::to(lmax));
      }
    }
    // face-halo cells: axis N needs F_N / L_N one layer outside the interior on both sides
    for (int f = tid; f < C::FACE_CELLS; f += NT) {
      const int n = f / C::FACES_PER_AXIS;
      const int ff = f - n * C::FACES_PER_AXIS;
      if (n == 0) eval_face<C, 0>(ff, npatch, qs, Fs, Ls, Rs, gather, tile * G, dt);
      else if (n == 1) eval_face<C, 1>(ff, npatch, qs, Fs, Ls, Rs, gather, tile * G, dt);
      else if constexpr (DIM == 3) eval_face<C, 2>(ff, npatch, qs, Fs, Ls, Rs, gather, tile * G, dt);
    }
    if (C::USE_TMA_STORE && tid == 0) tma_store_wait_read();   // previous tile's bulk store has drained `stage`
    __syncthreads();

    // ---------------------------------------------------------------- phase B: update, stage the result
    if (tid < npatch) {
      const Bits b = lam_bits[par * G + tid];
      lam_bits[par * G + tid] = 0;   // next written two tiles from now
      if (lambda_patch) lambda_patch[tile * G + tid] = FloatBits<T>::from(b);
      cta_lam = (b > cta_lam) ? b : cta_lam;
    }
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
      const int e = tid + r * NT;
      if (e < npatch * C::PD) {
        const CellIndex<C> ci = interior_cell<C>(e);
        const T dt_cell = gather.template step<C::GATHER>(dt, tile * G + ci.g);
        T qc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) qc[v] = q[r][v];
        // statement "Q_copy = Q_copy - 0.5*F[+1] + 0.5*F[-1]", axis by axis in order (test.cpp:60-77)
#pragma unroll
        for (int n = 0; n < DIM; ++n) {
          const int s = ci.g * C::SLOTS + ci.slot[n];
          const int ds = C::box_stride(n, n);
#pragma unroll
          for (int v = 0; v < NR; ++v) {
            const T* Fv = Fs + (n * NR + v) * FSTRIDE + s;
            qc[v] = Upd::flux(qc[v], Fv[ds], Fv[-ds]);
          }
        }
        // statement "Q_copy = 0.5*dt*(...) + Q_copy" from the original Q (test.cpp:78-95)
#pragma unroll
        for (int n = 0; n < DIM; ++n) {
          const int s = ci.g * C::SLOTS + ci.slot[n];
          const int ds = C::box_stride(n, n);
          const T l_plus = Ls[n * FSTRIDE + s + ds];
          const T l_minus = Ls[n * FSTRIDE + s - ds];
          const int dc = C::cell_stride(n);
#pragma unroll
          for (int v = 0; v < C::DV; ++v) {
            T q_plus, q_minus;
            if (C::STASH_Q) {
              const T* Rv = Rs + v * (G * C::NCELL) + ci.g * C::NCELL + ci.cell;
              q_plus = Rv[dc];
              q_minus = Rv[-dc];
            } else {
              const T* Qv = qs + (ci.g * C::NCELL + ci.cell) * NV + v;
              q_plus = Qv[dc * NV];
              q_minus = Qv[-dc * NV];
            }
            qc[v] = Upd::dissipation(qc[v], q[r][v], q_plus, q_minus, lam[r][n], l_plus, l_minus, dt_cell);
          }
        }
        fv_apply_source<Phys, Upd, T>(qc, q[r], dt_cell);     // "Q_copy = Q_copy + dt*S" (families with a source term)
        T* dst = stage + e * NV;
#pragma unroll
        for (int v = 0; v < NV; ++v) dst[v] = qc[v];
      }
    }
    if (C::USE_TMA_STORE) fence_proxy_async_smem();
    __syncthreads();

    // ---------------------------------------------------------------- phase C: staged interior -> HBM
    if (C::UNHALOED) {
      if (C::USE_TMA_STORE) {
        if (tid == 0) {
          if constexpr (!C::GATHER) {
            tma_store_1d(q_out + tile * (long long)C::OUT_ELEMS, stage,
                         (uint32_t)npatch * C::OUT_PATCH_ELEMS * (uint32_t)sizeof(T));
          } else {
            for (int g = 0; g < npatch; ++g)
              tma_store_1d(gather.q_out[tile * G + g], stage + g * C::OUT_PATCH_ELEMS,
                           C::OUT_PATCH_ELEMS * (uint32_t)sizeof(T));
          }
          tma_store_commit();
        }
      } else {
        for (int i = tid; i < npatch * C::OUT_PATCH_ELEMS; i += NT) {
          const int g = i / C::OUT_PATCH_ELEMS;
          gather.template out<C::GATHER>(q_out, tile * G + g, C::OUT_PATCH_ELEMS)[i - g * C::OUT_PATCH_ELEMS] = stage[i];
        }
      }
    } else {
      // haloed layout: interior rows are contiguous runs of P*NV values (test.cpp:96-104 writes all NV)
      constexpr int ROW = C::P * NV;
      for (int i = tid; i < npatch * C::OUT_PATCH_ELEMS; i += NT) {
        const int row = i / ROW;          // (patch, x_0 .. x_{DIM-2})
        const int col = i - row * ROW;
        int r = row;
        int cell = C::H;                  // first interior cell of the row along the fastest axis
#pragma unroll
        for (int m = DIM - 2; m >= 0; --m) {
          cell += (r % C::P + C::H) * C::cell_stride(m);
          r /= C::P;
        }
        gather.template out<C::GATHER>(q_out, tile * G + r, C::PATCH_ELEMS)[cell * NV + col] = stage[i];   // r is now the patch of the tile
      }
    }
  }

  if (C::USE_TMA_STORE && tid == 0) tma_store_wait_all();
  if (lambda_max != nullptr && tid < G && cta_lam != 0)
    atomicMax(reinterpret_cast<Bits*>(lambda_max), cta_lam);
}

// ------------------------------------------------------------------------------------------------
// host side: geometry + launch of one instantiation
struct FvLaunchInfo {
  int grid, block, smem_bytes, patches_per_tile, ctas_per_sm;
  int fused_allreduce = 0;   // 1: the kernel runs FvGather::peer's all-reduce(max) in its epilogue
};

template <class C>
struct FvLauncher {
  static cudaError_t prepare(FvLaunchInfo* info, long long n_patches) {
    static int cached_ctas_per_sm[64];
    static int cached_sms[64];
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (cached_ctas_per_sm[dev] == 0) {
      err = cudaFuncSetAttribute(fv_step_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
      if (err != cudaSuccess) return err;
      int per_sm = 0, sms = 0;
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fv_step_kernel<C>, C::NT, C::SMEM_BYTES);
      if (err != cudaSuccess) return err;
      err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (err != cudaSuccess) return err;
      if (per_sm < 1) return cudaErrorLaunchOutOfResources;
      cached_sms[dev] = sms;
      cached_ctas_per_sm[dev] = per_sm;
    }
    const long long n_tiles = (n_patches + C::G - 1) / C::G;
    const long long resident = (long long)cached_sms[dev] * cached_ctas_per_sm[dev];
    info->grid = (int)(n_tiles < resident ? n_tiles : resident);
    info->block = C::NT;
    info->smem_bytes = C::SMEM_BYTES;
    info->patches_per_tile = C::G;
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
    fv_step_kernel<C><<<info.grid, info.block, info.smem_bytes, stream>>>(
        static_cast<const T*>(q_in), static_cast<T*>(q_out), n_patches, static_cast<T>(dt),
        static_cast<T*>(lambda_patch), static_cast<T*>(lambda_max), make_gather<T>(gather));
    return cudaGetLastError();
  }
};

}  // namespace exahype
