// Mailbox protocol of the one-shot all-reduce(max) over NVLink peer memory (peer_reduce.cu), shared by the standalone
// reducer kernels and by patch kernels that run the exchange themselves (fv3d_pair_kernel.cuh).
//
// Every rank owns a mailbox [2][world] of 64-bit words, mapped into every peer through CUDA IPC.  A word carries the value
// (double or float bits; the reduced quantity is a maximum of absolute eigenvalues, so its sign bit is free) and, in bit
// 63, the epoch of the exchange it belongs to.  For exchange number `seq`, rank r stores its word into slot [seq & 1][r] of
// every peer's mailbox (its own included) with ONE relaxed 8-byte store -- value and flag cannot tear, no fence is needed
// on either side -- and a consumer polls its own slots [seq & 1][t] until their epoch bit is that of `seq`:
// (seq >> 1) & 1, which flips every time a slot is reused.  Two slots by sequence parity and one epoch bit are enough:
// a rank publishes exchange s+1 only after it has consumed exchange s, and it can consume s only after every peer has
// published s -- which a peer does after it has consumed s-1; so when a consumer looks for s, the slot holds either s
// or s-2 (the opposite epoch), never anything older.  Mailboxes start out as "exchange -1 / 0": slot 0 all zero bits,
// slot 1 with the epoch bit set (peer_reducer_create).
//
// Two ways to run one exchange:
//   blocking     publish, then wait for all peers, in the same place (stand-alone kernel; the last warp of a patch
//                kernel's grid: FV_PEER_BLOCKING).  The patch kernel's launch ends with the global maximum in *lambda_max.
//   split phase  (FV_PEER_LOOP, the device-resident time loop of SURVEY.md section 8e) the last warp of step k's grid
//                only publishes; the warps of step k+1 consume -- each reads the `world` mailbox slots of this device
//                and derives the same dt = cfl_dx / max -- right before their first use of dt, so that the wait for the
//                slowest peer overlaps the launch, the mbarrier set-up and the first ring fill of step k+1 instead of
//                extending step k.  max is exact and every rank divides the same two numbers: all ranks use identical dt.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace exahype {

struct PeerMail {
  unsigned long long word;   // bit 63: epoch of the exchange, bits 62..0: the value's bits (sign bit cleared)
};
constexpr unsigned long long kMailEpochBit = 1ull << 63;
__host__ __device__ inline unsigned long long mail_epoch(unsigned long long seq) { return ((seq >> 1) & 1ull) << 63; }

// who does what inside a patch kernel: nothing | epilogue publishes and waits | prologue consumes + epilogue publishes |
// prologue consumes only (the kernel has no exchange epilogue: a one-warp kernel behind it publishes)
enum { FV_PEER_NONE = 0, FV_PEER_BLOCKING = 1, FV_PEER_LOOP = 2, FV_PEER_CONSUME = 3 };

// exchange trace (debugging / attribution, profiles/): globaltimer stamps of exchange `seq` in trace[(seq % cap) * 8 + i]
enum {
  FV_TRACE_KERNEL_BEGIN = 0,   // first warp of the grid enters the kernel (loop mode)
  FV_TRACE_WAIT_BEGIN = 1,     // ... starts polling its mailbox for the exchange it consumes
  FV_TRACE_WAIT_END = 2,       // ... has every peer's value
  FV_TRACE_LAST_WARP = 3,      // last warp of the grid has taken its ticket
  FV_TRACE_PUBLISHED = 4,      // ... has stored the device maximum into every peer's mailbox
  FV_TRACE_ALL_SEEN = 5,       // blocking mode: ... has every peer's value
  FV_TRACE_WORDS = 8
};

// What a kernel needs to run the exchange itself.  Plain data: passed inside kernel parameters.
struct FvPeerFuse {
  PeerMail* const* peers = nullptr;   // device array: peer r's mailbox as mapped on this device (peers[rank] == mine)
  PeerMail* mine = nullptr;
  unsigned int* ticket = nullptr;     // device counter, zero between launches: the last warp to arrive runs the exchange
  int* error = nullptr;               // sticky, host-mapped: set to 1 by a timed-out wait
  unsigned long long seq = 0;         // exchange this launch publishes (blocking: and waits for)
  long long timeout_cycles = 0;
  int world = 0, rank = 0;
  int mode = FV_PEER_NONE;
  // device-resident time loop (FV_PEER_LOOP / FV_PEER_CONSUME)
  unsigned long long consume_seq = 0; // exchange whose maximum gives this launch's dt; 0: dt = *dt_in
  const void* dt_in = nullptr;        // device scalar of the kernel's T: the time step when there is nothing to consume
  void* dt_out = nullptr;             // device scalar: receives the dt this launch used (written by the grid's first warp)
  void* record = nullptr;             // device, four values of T: {dt used, global maximum consumed (0: none), this
                                      // device's maximum of this step, unused} -- the step's entry of the loop's history
  double cfl_dx = 0.0;                // CFL number x cell size
  unsigned long long* trace = nullptr;
  int trace_capacity = 0;
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void peer_trace(const FvPeerFuse& pf, unsigned long long seq, int what) {
  if (pf.trace != nullptr)
    pf.trace[(seq % (unsigned long long)pf.trace_capacity) * FV_TRACE_WORDS + what] = global_timer_ns();
}
// a timed-out wait: sticky flag in host-mapped memory, visible to the host without synchronising anything
__device__ __forceinline__ void peer_raise_timeout(int* error) {
  *reinterpret_cast<volatile int*>(error) = 1;
  __threadfence_system();
}

template <typename T> struct MailBits;
template <> struct MailBits<double> {
  static __device__ __forceinline__ unsigned long long to(double x) { return (unsigned long long)__double_as_longlong(x); }
  static __device__ __forceinline__ double from(unsigned long long b) { return __longlong_as_double((long long)b); }
  static __device__ __forceinline__ double poison() { return __longlong_as_double(0x7ff8000000000000ll); }
};
template <> struct MailBits<float> {
  static __device__ __forceinline__ unsigned long long to(float x) { return __float_as_uint(x); }
  static __device__ __forceinline__ float from(unsigned long long b) { return __uint_as_float((unsigned)b); }
  static __device__ __forceinline__ float poison() { return __uint_as_float(0x7fc00000u); }
};

// publish v as this rank's value of exchange `seq` in peer t's mailbox: one relaxed 8-byte store over NVLink
template <typename T>
__device__ __forceinline__ void peer_publish_to(PeerMail* const* peers, int world, int rank, int t, unsigned long long seq,
                                                T v) {
  PeerMail* dst = peers[t] + (int)(seq & 1ull) * world + rank;
  st_relaxed_sys(&dst->word, (MailBits<T>::to(v) & ~kMailEpochBit) | mail_epoch(seq));
}
// wait for peer t's value of exchange `seq` in this rank's mailbox; a timeout raises the sticky error flag and returns
// NaN, so that everything derived from the exchange is visibly poisoned instead of silently rank-local
template <typename T>
__device__ __forceinline__ T peer_wait_for(const PeerMail* mine, int world, int t, unsigned long long seq,
                                           long long timeout_cycles, int* error) {
  const PeerMail* src = mine + (int)(seq & 1ull) * world + t;
  const unsigned long long epoch = mail_epoch(seq);
  const long long t0 = clock64();
  unsigned long long w;
  while (((w = ld_relaxed_sys(&src->word)) & kMailEpochBit) != epoch) {
    if (clock64() - t0 > timeout_cycles) {
      peer_raise_timeout(error);
      return MailBits<T>::poison();
    }
  }
  return MailBits<T>::from(w & ~kMailEpochBit);
}
// One thread's share of a blocking exchange: publish v to peer t, wait for peer t's value.
template <typename T>
__device__ __forceinline__ T peer_exchange_with(PeerMail* const* peers, PeerMail* mine, int world, int rank, int t,
                                                unsigned long long seq, long long timeout_cycles, int* error, T v) {
  peer_publish_to<T>(peers, world, rank, t, seq, v);
  return peer_wait_for<T>(mine, world, t, seq, timeout_cycles, error);
}
// max over the lanes of a warp with std::max semantics; a NaN (timed-out wait) in any lane poisons the result
template <typename T>
__device__ __forceinline__ T peer_warp_max(T got) {
  const bool bad = __any_sync(0xffffffffu, got != got);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, got, o);
    got = (got < other) ? other : got;
  }
  return bad ? MailBits<T>::poison() : got;
}

// A whole warp waits for every peer's value of exchange `seq` and returns their maximum (NaN after a timeout).  The
// loop is warp-uniform on purpose (votes, no per-lane exits): patch kernels call this next to warp-uniform cursors that
// live in uniform registers, and a lane-divergent spin loop in their prologue cost 26 registers per thread.
template <typename T>
__device__ __forceinline__ T peer_wait_max(const PeerMail* mine, int world, int lane, unsigned long long seq,
                                           long long timeout_cycles, int* error) {
  T best = T(0);
  bool timed_out = false;
  for (int base = 0; base < world && !timed_out; base += 32) {
    const int t = (base + lane < world) ? base + lane : world - 1;
    const PeerMail* src = mine + (int)(seq & 1ull) * world + t;
    const unsigned long long epoch = mail_epoch(seq);
    const long long t0 = clock64();
    unsigned long long w;
    while (!__all_sync(0xffffffffu, ((w = ld_relaxed_sys(&src->word)) & kMailEpochBit) == epoch)) {
      if (__any_sync(0xffffffffu, clock64() - t0 > timeout_cycles)) { timed_out = true; break; }
    }
    if (!timed_out) {
      const T got = MailBits<T>::from(w & ~kMailEpochBit);
      best = (best < got) ? got : best;
    }
  }
  if (timed_out) {
    if (lane == 0) peer_raise_timeout(error);
    return MailBits<T>::poison();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, best, o);
    best = (best < other) ? other : best;
  }
  return best;
}

// The time step of this launch (all lanes of a warp call it together; `first_warp` marks the one warp of the grid that
// leaves the step's record behind).  With something to consume: the global maximum of exchange
// consume_seq -> dt = cfl_dx / max (a zero maximum puts no constraint on the step: dt stays *dt_in).  Otherwise
// *dt_in, or the kernel's by-value argument when there is no device scalar either.
template <typename T>
__device__ __forceinline__ T peer_loop_dt(const FvPeerFuse& pf, int lane, T dt_by_value, bool first_warp) {
  if (pf.dt_in == nullptr) return dt_by_value;
  T dt = *static_cast<const T*>(pf.dt_in);
  T lam = T(0);
  if (pf.consume_seq != 0) {
    if (first_warp && lane == 0) peer_trace(pf, pf.consume_seq, FV_TRACE_WAIT_BEGIN);
    lam = peer_wait_max<T>(pf.mine, pf.world, lane, pf.consume_seq, pf.timeout_cycles, pf.error);
    if (first_warp && lane == 0) peer_trace(pf, pf.consume_seq, FV_TRACE_WAIT_END);
    if (lam != lam) dt = lam;                        // timed out: poisoned
    else if (lam > T(0)) dt = static_cast<T>(pf.cfl_dx) / lam;
  }
  if (first_warp && lane == 0) {
    if (pf.dt_out != nullptr) *static_cast<T*>(pf.dt_out) = dt;
    if (pf.record != nullptr) {
      static_cast<T*>(pf.record)[0] = dt;
      static_cast<T*>(pf.record)[1] = lam;
    }
  }
  return dt;
}

// Epilogue of a patch kernel whose warps have all done atomicMax(lambda_max, their maximum): the last warp of the grid
// to arrive (ticket counter) runs this device's side of the exchange (world <= 32: one lane per peer).  Called by every
// warp of the grid, all lanes.
//   FV_PEER_BLOCKING  publish, wait for every peer, leave the global maximum in *lambda_max;
//   FV_PEER_LOOP      publish only (the next launch consumes); *lambda_max keeps this device's maximum.
template <typename T, typename BitsT>
__device__ __forceinline__ void fused_allreduce_max(const FvPeerFuse& pf, T* lambda_max, int lane, unsigned total_warps) {
  unsigned ticket = 0;
  if (lane == 0) {
    __threadfence();                                   // this warp's atomicMax before its ticket
    ticket = atomicAdd(pf.ticket, 1u);
  }
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket != total_warps - 1) return;
  __threadfence();                                     // every other warp's atomicMax is visible now
  if (lane == 0) peer_trace(pf, pf.seq, FV_TRACE_LAST_WARP);
  T v = T(0);
  if (pf.mode == FV_PEER_LOOP) {
    // the accumulator is the loop's own and is left at zero for the next launch: no memset between two steps
    if (lane == 0) v = MailBits<T>::from((unsigned long long)atomicExch(reinterpret_cast<BitsT*>(lambda_max), (BitsT)0));
    v = __shfl_sync(0xffffffffu, v, 0);
    if (lane < pf.world) peer_publish_to<T>(pf.peers, pf.world, pf.rank, lane, pf.seq, v);
    __syncwarp();
    if (lane == 0) {
      peer_trace(pf, pf.seq, FV_TRACE_PUBLISHED);
      if (pf.record != nullptr) static_cast<T*>(pf.record)[2] = v;
      *pf.ticket = 0;                                  // ready for the next launch (stream-ordered behind this one)
    }
    return;
  }
  if (lane == 0) v = MailBits<T>::from((unsigned long long)atomicMax(reinterpret_cast<BitsT*>(lambda_max), (BitsT)0));
  v = __shfl_sync(0xffffffffu, v, 0);
  T got = v;
  if (lane < pf.world) {
    peer_publish_to<T>(pf.peers, pf.world, pf.rank, lane, pf.seq, v);
    if (lane == 0) peer_trace(pf, pf.seq, FV_TRACE_PUBLISHED);
    got = peer_wait_for<T>(pf.mine, pf.world, lane, pf.seq, pf.timeout_cycles, pf.error);
  }
  got = peer_warp_max(got);
  if (lane == 0) {
    peer_trace(pf, pf.seq, FV_TRACE_ALL_SEEN);
    *lambda_max = got;
    *pf.ticket = 0;
  }
}

}  // namespace exahype
