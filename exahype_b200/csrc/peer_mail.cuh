// Mailbox protocol of the one-shot all-reduce(max) over NVLink peer memory (peer_reduce.cu), shared by the standalone
// reducer kernel and by patch kernels that run the same exchange in their epilogue (fv3d_pair_kernel.cuh).
//
// Every rank owns a mailbox [2][world] of (value bits, sequence number), mapped into every peer through CUDA IPC.  For
// exchange number `seq`, rank r stores (value, seq) into slot [seq & 1][r] of every peer's mailbox and spins on its own
// slots [seq & 1][t] until their sequence number is seq.  Two slots by sequence parity are enough: a rank cannot finish
// exchange s+1 before every peer has published s+1, which a peer only does after it has consumed exchange s.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace exahype {

struct PeerMail {
  unsigned long long bits;   // the value (double or float bits)
  unsigned long long seq;    // exchange number the value belongs to (0 = never written)
};

// What a kernel needs to run the exchange itself; world <= 1: no exchange.  Plain data: passed inside kernel parameters.
struct FvPeerFuse {
  PeerMail* const* peers = nullptr;   // device array: peer r's mailbox as mapped on this device (peers[rank] == mine)
  PeerMail* mine = nullptr;
  unsigned int* ticket = nullptr;     // device counter, zero between launches: the last warp to arrive runs the exchange
  int* error = nullptr;               // set to 1 by a timed-out wait
  unsigned long long seq = 0;
  long long timeout_cycles = 0;
  int world = 0, rank = 0;
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <typename T> struct MailBits;
template <> struct MailBits<double> {
  static __device__ __forceinline__ unsigned long long to(double x) { return (unsigned long long)__double_as_longlong(x); }
  static __device__ __forceinline__ double from(unsigned long long b) { return __longlong_as_double((long long)b); }
};
template <> struct MailBits<float> {
  static __device__ __forceinline__ unsigned long long to(float x) { return __float_as_uint(x); }
  static __device__ __forceinline__ float from(unsigned long long b) { return __uint_as_float((unsigned)b); }
};

// One thread's share of exchange `seq`: publish v to peer t, wait for peer t's value.  Returns v on a timeout (and
// raises the error flag).
template <typename T>
__device__ __forceinline__ T peer_exchange_with(PeerMail* const* peers, PeerMail* mine, int world, int rank, int t,
                                                unsigned long long seq, long long timeout_cycles, int* error, T v) {
  const int slot = (int)(seq & 1ull) * world;
  PeerMail* dst = peers[t] + slot + rank;
  st_relaxed_sys(&dst->bits, MailBits<T>::to(v));
  st_release_sys(&dst->seq, seq);                     // the value is visible before its sequence number
  const PeerMail* src = mine + slot + t;
  const long long t0 = clock64();
  while (ld_acquire_sys(&src->seq) != seq) {
    if (clock64() - t0 > timeout_cycles) {
      atomicExch(error, 1);
      return v;
    }
  }
  return MailBits<T>::from(ld_acquire_sys(&src->bits));
}

// Epilogue of a patch kernel whose warps have all done atomicMax(lambda_max, their maximum): the last warp of the grid
// to arrive (ticket counter) exchanges the device's maximum with all peers (world <= 32: one lane per peer) and leaves the
// global maximum in *lambda_max.  Called by every warp of the grid, all lanes.
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
  T v = T(0);
  if (lane == 0) v = MailBits<T>::from((unsigned long long)atomicMax(reinterpret_cast<BitsT*>(lambda_max), (BitsT)0));
  v = __shfl_sync(0xffffffffu, v, 0);
  T got = v;
  if (lane < pf.world)
    got = peer_exchange_with<T>(pf.peers, pf.mine, pf.world, pf.rank, lane, pf.seq, pf.timeout_cycles, pf.error, v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, got, o);
    got = (got < other) ? other : got;
  }
  if (lane == 0) {
    *lambda_max = got;
    *pf.ticket = 0;                                    // ready for the next launch (stream-ordered behind this one)
  }
}

}  // namespace exahype
