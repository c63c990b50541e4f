// sm_100a row-marching kernel for the 2-D batched stateless FV Rusanov patch update.
//
// Same arithmetic, statement order and results as fv_patch_kernel.cuh (reference "Unit test/test.cpp":11-104 with the
// loop ranges of exahype/printers/CPPPrinter.py:116-137), different data flow.  The thread-per-cell kernel stages the
// whole tile, F_n / L_n of both axes, a stash of Q and the output in shared memory and is bound by shared-memory
// wavefronts (profiles/r01_ncu_c2_cell.txt: 85 % L1 data pipe at 46 % DRAM).  Here a WARP owns 32/P patches and marches
// through their rows along axis 0 (the slow index `i`); nothing is staged:
//
//   * lane <-> (patch of the warp, interior column k).  A row of P cells of one patch is one contiguous run of the AoS
//     batch, a cell of 4 fp64 variables is exactly one 32-byte sector: every lane loads ITS cell straight from HBM into
//     registers with one 256-bit load (LDG.E.256, or 128-bit loads when the buffers are only 16-byte aligned) and stores
//     its updated cell with one 256-bit store.  Loads run PF (2 or 3) rows ahead of the row being consumed, and a bulk L2
//     prefetch (cp.async.bulk.prefetch.L2, one lane per patch, no registers) runs 2 KB ahead of those.
//   * the axis-0 stencil lives in registers: rolling window {r-2, r-1, r} of the cell state, F_0 and L_0 (ring of PF+2
//     rows, row loop unrolled by the ring size so every ring index is a compile-time constant).
//   * only F_1, L_1 and the DV dissipated variables of the row cross lanes, through a warp-private double-buffered
//     shared row (conflict-free SoA, one __syncwarp per row; no CTA barriers, no mbarriers, no named barriers).
//   * the 2*P face-halo cells of axis 1 of each patch (one layer left and right of every interior row) are evaluated
//     32 at a time (lane <-> (patch, row)) into a warp-private table that the edge lanes read instead of the shared
//     row: requested at the start, evaluated behind rows 0 and 1 of the march.  Halo corners are never touched (they
//     are not inputs).
//   * per-patch max eigenvalue: running maximum in registers, segmented warp-shuffle reduction at the end of the patch.
//
//   * ONE warp per CTA: a CTA's slot is handed to the next CTA only when its last warp is done (inst_euler2d.cu).
//
// Shared memory: COMPS * 128 values per warp (6 KB for Euler fp64 var0), 128 registers: 16 independent warps per SM, each
// streaming rows of 512-1024 contiguous bytes.  Measured (round 2): C2 0.206 ms = 0.89-0.90 of the measured HBM copy
// peak, C4 0.72 ms = 0.855 (0.65 ms = 0.94-0.96 when the un-haloed output does not repeat the auxiliary variable:
// UNKNOWNS_ONLY), C4 fp32 0.369 ms = 0.834.
#pragma once

#include "fv_patch_kernel.cuh"

#ifndef EXAHYPE_2D_WHATIF_NO_EXCHANGE
#define EXAHYPE_2D_WHATIF_NO_EXCHANGE 0   // 1: what-if tuning build, wrong results (see march_row)
#endif

namespace exahype {

template <class Phys_, class Upd_, typename T_, int P_, int H_, int WPC_, int MINB_, bool DISS_ALL_, bool UNHALOED_,
          int VEC_, int PF_ = 2, bool GATHER_ = false, bool UNKNOWNS_ONLY_ = false>
struct Fv2dMarchConfig {
  using Phys = Phys_;
  using Upd = Upd_;
  using T = T_;
  static constexpr int DIM = 2, P = P_, H = H_, WPC = WPC_, MINB = MINB_, VEC = VEC_;
  static constexpr int PF = PF_;                          // rows between a register load and its use
  static constexpr int RING = PF + 2;                     // rows r-1, r, r+1 .. r+PF live in registers
  static_assert(PF >= 1 && PF <= 4, "register prefetch distance");
  static constexpr bool DISS_ALL = DISS_ALL_, UNHALOED = UNHALOED_, GATHER = GATHER_;
  static_assert(P >= 1 && P <= 32 && 32 % P == 0, "row marching needs a patch side that divides the warp");
  static_assert(H >= 1 && WPC >= 1 && WPC <= 32, "march geometry");

  static constexpr int NR = Phys::NR, NA = Phys::NA, NV = NR + NA;
  static constexpr int S = P + 2 * H;
  static constexpr int PPW = 32 / P;                      // patches per warp
  static constexpr int NROW = P + 2;                      // rows a patch needs: one halo layer each side
  static constexpr int CELL_BYTES = NV * (int)sizeof(T);
  static constexpr int PATCH_ELEMS = S * S * NV;
  // EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY: un-haloed output without the auxiliary variables, [P][P][NR].  The step never
  // changes them, and copying them through is 13 % of the traffic of a 32x32 shallow-water batch (ncu: 4.73 GB moved
  // against 4.03 GB algorithmic, at the HBM wall).
  static constexpr bool UNKNOWNS_ONLY = UNKNOWNS_ONLY_ && UNHALOED_ && (NA > 0);
  static constexpr int OUT_NV = UNKNOWNS_ONLY ? NR : NV;
  static constexpr int OUT_PATCH_ELEMS = P * P * OUT_NV;
  static constexpr int DV = DISS_ALL ? NR : 1;
  static constexpr int COMPS = NR + 1 + DV;               // F_1[NR], L_1, Q[DV] cross lanes
  static constexpr int NT = WPC * 32;
  static constexpr int ROW_BYTES = S * CELL_BYTES;        // one haloed row of a patch (contiguous in the AoS batch)
  // rows are pulled into L2 ahead of the register loads by bulk prefetches (one lane per patch, no registers).  The
  // window is short on purpose: at 6 TB/s some 80 MB stream through the 126 MB L2 every 13 us, so rows requested a
  // whole patch ahead are evicted again before the march reaches them (measured: +11 % DRAM reads on 32x32 patches).
#ifndef EXAHYPE_2D_L2
#define EXAHYPE_2D_L2 1
#endif
#ifndef EXAHYPE_2D_L2_BYTES
#define EXAHYPE_2D_L2_BYTES 2048
#endif
  // Small patches (all their rows within EXAHYPE_2D_L2_WHOLE bytes: 16x16 fp64 is 10 KB) are requested whole by one
  // prefetch before the march: no per-row prefetch instructions at all (C2 0.2152 -> 0.2136 ms); the eviction argument
  // above concerns the 37 KB patches, which keep the short window.
  // (one request per row: 2 / 4 / 8 rows per request, or a 4 KB window, changed nothing or lost 1-3 % on 32x32 patches)
#ifndef EXAHYPE_2D_L2_WHOLE
#define EXAHYPE_2D_L2_WHOLE 12288
#endif
  static constexpr bool L2_BULK = EXAHYPE_2D_L2 && (ROW_BYTES % 16 == 0);
  static constexpr int L2_ROWS_WANTED = (NROW * ROW_BYTES <= EXAHYPE_2D_L2_WHOLE)
                                            ? NROW : PF + 1 + (EXAHYPE_2D_L2_BYTES + ROW_BYTES - 1) / ROW_BYTES;
  static constexpr int L2_ROWS = L2_ROWS_WANTED < NROW ? L2_ROWS_WANTED : NROW;
  static_assert(VEC == 32 || VEC == 16 || VEC == (int)sizeof(T), "vector width of the global accesses");
  static_assert(VEC == (int)sizeof(T) || CELL_BYTES % VEC == 0, "a cell must be a whole number of vectors");

  // warp-private exchange area: [XVEC][XS] vectors of XPV values + [XREM][XS] values (below); slots [0,32) row buffer 0, [32,64) row buffer 1,
  // [64,96) left face-halo table (patch of the warp, row), [96,128) right face-halo table
  // CellData form: the patches come through per-patch pointers whose alignment the launcher cannot see (the pointer
  // arrays live in device memory).  Each lane looks at ITS patch's pointers and takes the 256-bit access when they
  // are 32-byte aligned, the two 128-bit ones otherwise (per-lane branch, divergent only inside a warp whose patches
  // differ): gathered C2 on aligned patches 0.272 -> see profiles/README.md.
#ifndef EXAHYPE_2D_GATHER_DYN_WIDE
#define EXAHYPE_2D_GATHER_DYN_WIDE 1
#endif
  static constexpr bool DYN_WIDE = EXAHYPE_2D_GATHER_DYN_WIDE && GATHER && VEC == 16 && sizeof(T) == 8 && CELL_BYTES % 32 == 0;
  static constexpr int XS = 128;
  // The COMPS values a lane publishes are packed into 16-byte vectors, [XVEC][XS] vectors per warp: a neighbour's F_1,
  // L_1, Q arrive with XVEC 128-bit shared loads instead of COMPS scalar ones (C2: 3 instead of 6 per side and row; a
  // 64-bit load is two wavefronts per warp plus two more when the edge lane's table entry shares a bank with an
  // interior reader -- 15 rows of 16 --, a 128-bit load four plus one: 24 -> 15 wavefronts per side and row).
#ifndef EXAHYPE_2D_XVEC
#define EXAHYPE_2D_XVEC 1
#endif
  static constexpr int XPV = EXAHYPE_2D_XVEC ? 16 / (int)sizeof(T) : 1;       // values per exchange vector
  static constexpr int XVEC = COMPS / XPV;                                    // whole vectors; the rest stays scalar:
  static constexpr int XREM = COMPS - XVEC * XPV;                             // [XVEC][XS] vectors, then [XREM][XS] values
  static constexpr int WARP_BYTES = COMPS * XS * (int)sizeof(T);
  static constexpr int SMEM_BYTES = WPC * WARP_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "exchange area does not fit");
};

// widest global access a cell allows: 256-bit (one 32-byte sector per 4-variable fp64 cell), else 128-bit, else scalar
template <typename T, int NV>
struct Fv2dVec {
  static constexpr int BYTES = NV * (int)sizeof(T);
  static constexpr int NARROW = (BYTES % 16 == 0) ? 16 : (int)sizeof(T);
  static constexpr int WIDE = (BYTES % 32 == 0) ? 32 : NARROW;
};

// one cell (NV values) between global memory and registers, VEC bytes per instruction
template <class C>
__device__ __forceinline__ void load_cell(const typename C::T* p, typename C::T (&q)[C::NV]) {
  using T = typename C::T;
  if constexpr (C::VEC == 32 && sizeof(T) == 8) {
#pragma unroll
    for (int v = 0; v < C::NV; v += 4)
      asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                   : "=d"(q[v]), "=d"(q[v + 1]), "=d"(q[v + 2]), "=d"(q[v + 3]) : "l"(p + v));
  } else if constexpr (C::VEC == 32 && sizeof(T) == 4) {
#pragma unroll
    for (int v = 0; v < C::NV; v += 8)
      asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(q[v]), "=f"(q[v + 1]), "=f"(q[v + 2]), "=f"(q[v + 3]), "=f"(q[v + 4]), "=f"(q[v + 5]),
                     "=f"(q[v + 6]), "=f"(q[v + 7]) : "l"(p + v));
  } else if constexpr (C::VEC == 16 && sizeof(T) == 8) {
#pragma unroll
    for (int v = 0; v < C::NV; v += 2)
      asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(q[v]), "=d"(q[v + 1]) : "l"(p + v));
  } else if constexpr (C::VEC == 16 && sizeof(T) == 4) {
#pragma unroll
    for (int v = 0; v < C::NV; v += 4)
      asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(q[v]), "=f"(q[v + 1]), "=f"(q[v + 2]), "=f"(q[v + 3]) : "l"(p + v));
  } else {
#pragma unroll
    for (int v = 0; v < C::NV; ++v) q[v] = *reinterpret_cast<const volatile T*>(p + v);
  }
}

template <class C>
__device__ __forceinline__ void store_cell(typename C::T* p, const typename C::T (&q)[C::NV]) {
  using T = typename C::T;
  if constexpr (C::VEC == 32 && sizeof(T) == 8) {
#pragma unroll
    for (int v = 0; v < C::NV; v += 4)
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + v), "d"(q[v]), "d"(q[v + 1]), "d"(q[v + 2]),
                   "d"(q[v + 3]) : "memory");
  } else if constexpr (C::VEC == 32 && sizeof(T) == 4) {
#pragma unroll
    for (int v = 0; v < C::NV; v += 8)
      asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + v), "f"(q[v]), "f"(q[v + 1]),
                   "f"(q[v + 2]), "f"(q[v + 3]), "f"(q[v + 4]), "f"(q[v + 5]), "f"(q[v + 6]), "f"(q[v + 7]) : "memory");
  } else if constexpr (C::VEC == 16 && sizeof(T) == 8) {
#pragma unroll
    for (int v = 0; v < C::NV; v += 2)
      asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p + v), "d"(q[v]), "d"(q[v + 1]) : "memory");
  } else if constexpr (C::VEC == 16 && sizeof(T) == 4) {
#pragma unroll
    for (int v = 0; v < C::NV; v += 4)
      asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + v), "f"(q[v]), "f"(q[v + 1]), "f"(q[v + 2]),
                   "f"(q[v + 3]) : "memory");
  } else {
#pragma unroll
    for (int v = 0; v < C::NV; ++v) p[v] = q[v];
  }
}

// the NR unknowns of a cell into an output without auxiliary variables (cells of NR values: 24 bytes for fp64 shallow
// water -- lanes write consecutive cells, every sector is completed inside L2 before it is written back: C4 0.718 ->
// 0.652 ms).  Packing the row into whole 32-byte vectors through a warp-private staging row first was slower (0.683 ms:
// three shared stores, a __syncwarp and two shared loads per row cost more than the partial sectors).  fp32 cells of 12
// bytes in 4-byte pieces do not pay at all (C4 fp32 0.400 -> 0.416 ms), either way.
template <class C>
__device__ __forceinline__ void store_unknowns(typename C::T* p, const typename C::T (&q)[C::NV]) {
#pragma unroll
  for (int v = 0; v < C::NR; ++v) p[v] = q[v];
}

// CellData form (C::DYN_WIDE): the vector width follows the alignment of this lane's patch
template <class C>
__device__ __forceinline__ void load_cell(const typename C::T* p, typename C::T (&q)[C::NV], bool wide) {
  if constexpr (C::DYN_WIDE) {
    if (wide) {
#pragma unroll
      for (int v = 0; v < C::NV; v += 4)
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(q[v]), "=d"(q[v + 1]), "=d"(q[v + 2]), "=d"(q[v + 3]) : "l"(p + v));
      return;
    }
  }
  load_cell<C>(p, q);
}
template <class C>
__device__ __forceinline__ void store_cell(typename C::T* p, const typename C::T (&q)[C::NV], bool wide) {
  if constexpr (C::DYN_WIDE) {
    if (wide) {
#pragma unroll
      for (int v = 0; v < C::NV; v += 4)
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + v), "d"(q[v]), "d"(q[v + 1]), "d"(q[v + 2]),
                     "d"(q[v + 3]) : "memory");
      return;
    }
  }
  store_cell<C>(p, q);
}

__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One slot of the exchange area: the COMPS values of a cell -- F_1[NR], L_1, Q[DV] -- as XVEC vectors of XPV values
// ([XVEC][XS] vectors per warp) followed by XREM scalars ([XREM][XS] values)
template <class C>
__device__ __forceinline__ void exchange_store(typename C::T* X, int slot, const typename C::T (&v)[C::COMPS]) {
  using T = typename C::T;
#pragma unroll
  for (int w = 0; w < C::XVEC; ++w) {
    T* p = X + ((long long)w * C::XS + slot) * C::XPV;
    if constexpr (C::XPV == 1) p[0] = v[w];
    else if constexpr (sizeof(T) == 8) *reinterpret_cast<double2*>(p) = make_double2(v[2 * w], v[2 * w + 1]);
    else *reinterpret_cast<float4*>(p) = make_float4(v[4 * w], v[4 * w + 1], v[4 * w + 2], v[4 * w + 3]);
  }
#pragma unroll
  for (int c = 0; c < C::XREM; ++c) X[(long long)(C::XVEC * C::XPV + c) * C::XS + slot] = v[C::XVEC * C::XPV + c];
}
template <class C>
__device__ __forceinline__ void exchange_load(const typename C::T* X, int slot, typename C::T (&v)[C::COMPS]) {
  using T = typename C::T;
#pragma unroll
  for (int w = 0; w < C::XVEC; ++w) {
    const T* p = X + ((long long)w * C::XS + slot) * C::XPV;
    if constexpr (C::XPV == 1) v[w] = p[0];
    else if constexpr (sizeof(T) == 8) {
      const double2 a = *reinterpret_cast<const double2*>(p);
      v[2 * w] = a.x; v[2 * w + 1] = a.y;
    } else {
      const float4 a = *reinterpret_cast<const float4*>(p);
      v[4 * w] = a.x; v[4 * w + 1] = a.y; v[4 * w + 2] = a.z; v[4 * w + 3] = a.w;
    }
  }
#pragma unroll
  for (int c = 0; c < C::XREM; ++c) v[C::XVEC * C::XPV + c] = X[(long long)(C::XVEC * C::XPV + c) * C::XS + slot];
}
// what a cell publishes: F_1[NR], L_1, Q[DV]
template <class C>
__device__ __forceinline__ void exchange_pack(typename C::T (&v)[C::COMPS], const typename C::T (&f1)[C::NR],
                                              typename C::T l1, const typename C::T (&q)[C::NV]) {
#pragma unroll
  for (int c = 0; c < C::NR; ++c) v[c] = f1[c];
  v[C::NR] = l1;
#pragma unroll
  for (int c = 0; c < C::DV; ++c) v[C::NR + 1 + c] = q[c];
}

// Per-lane state of the march.  Ring index = row % RING (compile-time in the unrolled loop).
template <class C>
struct RowMarch {
  using T = typename C::T;
  const T* row_ptr;          // this lane's cell in marching row 0 (haloed i = H-1, j = k+H) of its patch
  const unsigned char* l2_ptr;   // marching row 0, haloed column 0 of the patch (lanes with k == 0 prefetch)
  T* out_ptr;                // this lane's cell in interior row 0 of the output
  T* X;                      // warp-private exchange area: [XVEC][XS] vectors, [XREM][XS] values
  T dt;
  int lane, k;
  int halo_slot_left, halo_slot_right;   // 64 + sub*P, 96 + sub*P (+ interior row)
  bool store_ok;
  bool wide_in, wide_out;    // C::DYN_WIDE: this lane's patch is 32-byte aligned (input / output)
};

// One row of the march.  SLOT = r & 3 at compile time; r itself is a run-time (warp-uniform) value.
//   r = 0        halo row:  F_0, L_0 only
//   r = 1        first interior row: publish F_1 / L_1 / Q for the neighbours, nothing to update yet
//   r = 2..P     publish, and update row r-1 (needs F_0 of rows r-2 and r, neighbours of row r-1 published last step)
//   r = P+1      halo row:  F_0, L_0, update row P
template <class C, int SLOT>
__device__ __forceinline__ void march_row(const RowMarch<C>& m, int r, typename C::T (&q)[C::RING][C::NV],
                                          typename C::T (&f0)[C::RING][C::NR], typename C::T (&l0)[C::RING],
                                          typename C::T& l1_mid, typename C::T (&q_old)[C::DV],
                                          typename C::T& lam_local) {
  using T = typename C::T;
  using Phys = typename C::Phys;
  using Upd = typename C::Upd;
  constexpr int NV = C::NV, NR = C::NR, DV = C::DV, XS = C::XS, P = C::P;
  constexpr int RING = C::RING, PF = C::PF;
  constexpr int NEW = SLOT, MID = (SLOT + RING - 1) % RING, OLD = (SLOT + RING - 2) % RING, PRE = (SLOT + PF) % RING;
  static_assert(PRE == OLD, "the prefetched row takes the ring slot of row r-2");

  // row r+PF -> the ring slot that held row r-2 (whose dissipated variables were saved to q_old last step)
  if (r + PF < C::NROW) load_cell<C>(m.row_ptr + (long long)(r + PF) * (C::S * NV), q[PRE], m.wide_in);
  if (C::L2_BULK && C::L2_ROWS < C::NROW && m.k == 0 && r + C::L2_ROWS < C::NROW)
    l2_prefetch_bulk(m.l2_ptr + (long long)(r + C::L2_ROWS) * C::ROW_BYTES, C::ROW_BYTES);

  const auto pr = Phys::template prims<T>(q[NEW]);
  Phys::template flux<0, T>(q[NEW], pr, f0[NEW]);
  l0[NEW] = Phys::template eigen<0, T>(q[NEW], pr);

  const bool inner = (r >= 1) && (r <= P);
  T f1[NR], l1_new = T(0);
  if (inner) {
    Phys::template flux<1, T>(q[NEW], pr, f1);
    l1_new = Phys::template eigen<1, T>(q[NEW], pr);
    lam_local = fv_max(lam_local, fv_max(l0[NEW], l1_new));
  }

  if (r >= 2) {
    // ---------------------------------------------------------------- update interior row r-1
    const int rb = (r - 1) & 1;
    const int sl = (m.k == 0) ? (m.halo_slot_left + (r - 2)) : (rb * 32 + m.lane - 1);
    const int sr = (m.k == P - 1) ? (m.halo_slot_right + (r - 2)) : (rb * 32 + m.lane + 1);
    T xl[C::COMPS], xr[C::COMPS];      // the neighbours' F_1[NR], L_1, Q[DV]
#if !EXAHYPE_2D_WHATIF_NO_EXCHANGE
    exchange_load<C>(m.X, sl, xl);
    exchange_load<C>(m.X, sr, xr);
#endif
    T qc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) qc[v] = q[MID][v];
    // "Q_copy = Q_copy - 0.5*F[+1] + 0.5*F[-1]" for axis 0 then axis 1 (test.cpp:60-77)
#pragma unroll
    for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], f0[NEW][v], f0[OLD][v]);
#pragma unroll
#if EXAHYPE_2D_WHATIF_NO_EXCHANGE   // what-if build (WRONG results): how fast would the march be without the lane exchange?
    for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], f0[MID][v], f0[NEW][v]);
#else
    for (int v = 0; v < NR; ++v) qc[v] = Upd::flux(qc[v], xr[v], xl[v]);
#endif
    // "Q_copy = 0.5*dt*(...) + Q_copy" from the original Q, axis 0 then axis 1 (test.cpp:78-95)
#pragma unroll
    for (int v = 0; v < DV; ++v)
      qc[v] = Upd::dissipation(qc[v], q[MID][v], q[NEW][v], q_old[v], l0[MID], l0[NEW], l0[OLD], m.dt);
#if EXAHYPE_2D_WHATIF_NO_EXCHANGE
    for (int v = 0; v < DV; ++v)
      qc[v] = Upd::dissipation(qc[v], q[MID][v], q[NEW][v], q_old[v], l1_mid, l0[NEW], l0[OLD], m.dt);
#else
    {
      const T l_plus = xr[NR], l_minus = xl[NR];
#pragma unroll
      for (int v = 0; v < DV; ++v)
        qc[v] = Upd::dissipation(qc[v], q[MID][v], xr[NR + 1 + v], xl[NR + 1 + v], l1_mid, l_plus, l_minus,
                                 m.dt);
    }
#endif
    fv_apply_source<Phys, Upd, T>(qc, q[MID], m.dt);          // "Q_copy = Q_copy + dt*S" (families with a source term)
    if constexpr (C::UNKNOWNS_ONLY) {
      if (m.store_ok) store_unknowns<C>(m.out_ptr + (long long)(r - 2) * (P * NR), qc);
    } else {
      if (m.store_ok) store_cell<C>(m.out_ptr + (long long)(r - 2) * ((C::UNHALOED ? P : C::S) * NV), qc, m.wide_out);
    }
  }

  // row r-1 becomes row r-2 of the next step: keep what the dissipation needs of it before its ring slot is reloaded
#pragma unroll
  for (int v = 0; v < DV; ++v) q_old[v] = q[MID][v];
  l1_mid = l1_new;

#if EXAHYPE_2D_WHATIF_NO_EXCHANGE
  if (inner) {
    T keep = l1_new;
    for (int v = 0; v < NR; ++v) keep = fv_max(keep, f1[v]);
    lam_local = fv_max(lam_local, keep * T(1e-30));      // keep F_1 alive
  }
  return;
#endif
  if (inner) {
    // ---------------------------------------------------------------- publish row r for the neighbouring lanes
    T pub[C::COMPS];
    exchange_pack<C>(pub, f1, l1_new, q[NEW]);
    exchange_store<C>(m.X, (r & 1) * 32 + m.lane, pub);
  }
  __syncwarp();
}

// RING consecutive rows, one body per ring slot; returns true when the patch is finished
template <class C, int SLOT, class... A>
__device__ __forceinline__ bool march_ring(const RowMarch<C>& m, int& r, A&... a) {
  march_row<C, SLOT>(m, r, a...);
  if (++r >= C::NROW) return true;
  if constexpr (SLOT + 1 < C::RING) return march_ring<C, SLOT + 1>(m, r, a...);
  else return false;
}

// One warp = one unit of PPW patches; the grid covers the batch and the hardware hands a finished CTA's slot to the next
// one.  Order inside a unit: L2 prefetch, rows 0..PF-1, face-halo cells requested; rows 0 and 1 of the march (they need
// nothing but the rows already there); face-halo table; the rest of the march.
//
// Measured and dropped (profiles/r02_2d_persistent_ab.txt, all bitwise equal): a PERSISTENT grid of resident warps --
// units dealt round robin (C2 0.205 -> 0.241 ms) or claimed from a counter (0.228 ms), with or without the march's
// last rows loading the next unit's first rows into the free ring slots (0.250 / 0.241 ms: pipelining across units
// made it slower still), or asking for the next unit's head by L2 prefetch.  A fresh CTA that requests its whole patch
// with ONE bulk prefetch and then sits out the DRAM latency (14 % of all warp samples) leaves the memory system with
// long contiguous requests; these kernels run at 85-90 % of the copy peak, where the request pattern is worth more than
// the hidden latency.
template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
fv2d_march_kernel(const typename C::T* q_in, typename C::T* q_out, long long n_patches, typename C::T dt,
                  typename C::T* __restrict__ lambda_patch, typename C::T* __restrict__ lambda_max,
                  const FvGather<typename C::T> gather) {
  using T = typename C::T;
  using Phys = typename C::Phys;
  using Bits = typename FloatBits<T>::type;
  constexpr int P = C::P, H = C::H, S = C::S, NV = C::NV, NR = C::NR, DV = C::DV, XS = C::XS, PPW = C::PPW;
  static_assert(C::NROW > C::RING && C::RING >= 3, "the march is longer than its ring");

  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long unit = (long long)blockIdx.x * C::WPC + warp;     // one unit = PPW consecutive patches
  const long long first_patch = unit * PPW;
  if (first_patch >= n_patches) return;                             // whole warp leaves together
  // device-resident time step (peer_mail.cuh): every warp derives the same dt from this device's mailbox
  dt = peer_loop_dt<T>(gather.peer, lane, dt, unit == 0);

  RowMarch<C> m;
  m.X = reinterpret_cast<T*>(smem + warp * C::WARP_BYTES);
  m.lane = lane;
  const int sub = lane / P;
  m.k = lane - sub * P;
  m.halo_slot_left = 64 + sub * P;
  m.halo_slot_right = 96 + sub * P;
  // a ragged last unit recomputes the batch's last patch in its surplus lanes and stores nothing for them
  long long patch = first_patch + sub;
  m.store_ok = patch < n_patches;
  if (!m.store_ok) patch = n_patches - 1;
  // CellData form: this lane's patch through its own pointers, with its own time step
  const T* const patch_in = gather.template in<C::GATHER>(q_in, patch, C::PATCH_ELEMS);
  m.dt = gather.template step<C::GATHER>(dt, patch);
  m.row_ptr = patch_in + ((long long)(H - 1) * S + (m.k + H)) * NV;
  m.out_ptr = C::UNHALOED ? gather.template out<C::GATHER>(q_out, patch, C::OUT_PATCH_ELEMS) + m.k * C::OUT_NV
                          : gather.template out<C::GATHER>(q_out, patch, C::PATCH_ELEMS) + ((long long)H * S + (m.k + H)) * NV;
  m.l2_ptr = reinterpret_cast<const unsigned char*>(patch_in + (long long)(H - 1) * S * NV);
  // cells are 32 bytes and rows a whole number of cells: a 32-byte aligned patch has 32-byte aligned cells throughout
  m.wide_in = C::DYN_WIDE && (reinterpret_cast<uintptr_t>(patch_in) & 31) == 0;
  m.wide_out = C::DYN_WIDE && (reinterpret_cast<uintptr_t>(m.out_ptr) & 31) == 0;
  if (C::L2_BULK && m.k == 0) l2_prefetch_bulk(m.l2_ptr, C::L2_ROWS * C::ROW_BYTES);

  T q[C::RING][NV], f0[C::RING][NR], l0[C::RING], l1_mid = T(0), q_old[DV], lam_local = T(0);
#pragma unroll
  for (int w = 0; w < C::RING; ++w) {
#pragma unroll
    for (int v = 0; v < NV; ++v) q[w][v] = T(0);
#pragma unroll
    for (int v = 0; v < NR; ++v) f0[w][v] = T(0);
    l0[w] = T(0);
  }
#pragma unroll
  for (int v = 0; v < DV; ++v) q_old[v] = T(0);

  // rows 0..PF-1 of the march are requested first, the face-halo cells (lane <-> (patch, interior row)) behind them
#pragma unroll
  for (int w = 0; w < C::PF; ++w) load_cell<C>(m.row_ptr + w * (S * NV), q[w], m.wide_in);
  const T* left = patch_in + ((long long)(m.k + H) * S + (H - 1)) * NV;       // row = m.k of this lane's own patch
  T ql[NV], qr[NV];
  load_cell<C>(left, ql, m.wide_in);
  load_cell<C>(left + (P + 1) * NV, qr, m.wide_in);

  // ------------------------------------------------------------------ the march: rows 0..P+1, ring index = row % RING.
  // Rows 0 and 1 come before the face-halo table (first read at row 2): C2 0.2107 -> 0.2065 ms.
  march_row<C, 0>(m, 0, q, f0, l0, l1_mid, q_old, lam_local);
  march_row<C, 1>(m, 1, q, f0, l0, l1_mid, q_old, lam_local);
  {
    // ---------------------------------------------------------------- face-halo table
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const T(&qh)[NV] = side ? qr : ql;
      const auto pr = Phys::template prims<T>(qh);
      T F[NR];
      Phys::template flux<1, T>(qh, pr, F);
      T pub[C::COMPS];
      exchange_pack<C>(pub, F, Phys::template eigen<1, T>(qh, pr), qh);
      exchange_store<C>(m.X, (side ? 96 : 64) + lane, pub);
    }
    __syncwarp();
  }
  int r = 2;
  if (!march_ring<C, 2>(m, r, q, f0, l0, l1_mid, q_old, lam_local))
    while (!march_ring<C, 0>(m, r, q, f0, l0, l1_mid, q_old, lam_local)) {}

  // ------------------------------------------------------------------ max eigenvalue of the input state
  T lam = lam_local;
#pragma unroll
  for (int o = P / 2; o > 0; o >>= 1) lam = fv_max(lam, __shfl_xor_sync(0xffffffffu, lam, o));
  if (lambda_patch != nullptr && m.k == 0 && m.store_ok) lambda_patch[patch] = lam;
  if (lambda_max != nullptr) {
#pragma unroll
    for (int o = 16; o >= P && o > 0; o >>= 1) lam = fv_max(lam, __shfl_xor_sync(0xffffffffu, lam, o));
    if (lane == 0) atomicMax(reinterpret_cast<Bits*>(lambda_max), FloatBits<T>::to(lam));
  }
}

template <class C>
struct Fv2dMarchLauncher {
  static cudaError_t prepare(FvLaunchInfo* info, long long n_patches) {
    static int cached_ctas_per_sm[64];
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (cached_ctas_per_sm[dev] == 0) {
      err = cudaFuncSetAttribute(fv2d_march_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
      if (err != cudaSuccess) return err;
      int per_sm = 0;
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fv2d_march_kernel<C>, C::NT, C::SMEM_BYTES);
      if (err != cudaSuccess) return err;
      if (per_sm < 1) return cudaErrorLaunchOutOfResources;
      cached_ctas_per_sm[dev] = per_sm;
    }
    const long long units = (n_patches + C::PPW - 1) / C::PPW;
    if ((units + C::WPC - 1) / C::WPC > 0x7fffffffll) return cudaErrorInvalidValue;     // more CTAs than a grid holds
    info->grid = (int)((units + C::WPC - 1) / C::WPC);
    info->block = C::NT;
    info->smem_bytes = C::SMEM_BYTES;
    info->patches_per_tile = C::PPW * C::WPC;
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
    fv2d_march_kernel<C><<<info.grid, info.block, info.smem_bytes, stream>>>(
        static_cast<const T*>(q_in), static_cast<T*>(q_out), n_patches, static_cast<T>(dt),
        static_cast<T*>(lambda_patch), static_cast<T*>(lambda_max), make_gather<T>(gather));
    return cudaGetLastError();
  }
};

// picks the widest vector the buffers' alignment allows: 32-byte cells on 32-byte aligned buffers -> 256-bit accesses
template <class C32, class C16>
struct Fv2dMarchDispatch {
  static bool wide(const void* a, const void* b) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 31) == 0;
  }
  static cudaError_t prepare(FvLaunchInfo* info, long long n_patches) {
    return Fv2dMarchLauncher<C32>::prepare(info, n_patches);
  }
  static cudaError_t launch(const void* q_in, void* q_out, long long n_patches, double dt, void* lambda_patch,
                            void* lambda_max, cudaStream_t stream, const FvGatherRaw* gather = nullptr) {
    if (wide(q_in, q_out)) return Fv2dMarchLauncher<C32>::launch(q_in, q_out, n_patches, dt, lambda_patch, lambda_max, stream, gather);
    return Fv2dMarchLauncher<C16>::launch(q_in, q_out, n_patches, dt, lambda_patch, lambda_max, stream, gather);
  }
};

// what a generated unit (exahype.printers.CUDAPrinter) instantiates: one warp per CTA (inst_euler2d.cu), dispatch on buffer alignment
template <class Phys, class Upd, typename T, int P, int H, bool DA, bool UH>
using Fv2dMarchAuto =
    Fv2dMarchDispatch<Fv2dMarchConfig<Phys, Upd, T, P, H, 1, 16, DA, UH, Fv2dVec<T, Phys::NR + Phys::NA>::WIDE>,
                      Fv2dMarchConfig<Phys, Upd, T, P, H, 1, 16, DA, UH, Fv2dVec<T, Phys::NR + Phys::NA>::NARROW>>;
// CellData form: gathered patches are only known to be 16-byte aligned individually -> narrow accesses
template <class Phys, class Upd, typename T, int P, int H, bool DA, bool UH, int PF = 2>
using Fv2dMarchGather =
    Fv2dMarchLauncher<Fv2dMarchConfig<Phys, Upd, T, P, H, 1, 16, DA, UH, Fv2dVec<T, Phys::NR + Phys::NA>::NARROW, PF, true>>;

}  // namespace exahype
