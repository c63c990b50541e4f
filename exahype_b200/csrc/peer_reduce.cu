// One-shot all-reduce(max) of the admissible-time-step scalar over NVLink peer memory.
//
// The reference has no distributed code; the exchange this path needs is ONE scalar per step, so
// the collective is pure latency.  ncclAllReduce of 8 bytes costs ~17 us per step on 2 GPUs; here every rank owns a
// mailbox [2][world] in device memory, opened by every peer through CUDA IPC.  One tiny kernel per step, stream-ordered
// behind the patch-update kernel:
//     thread t:  store (value, seq) into peer t's mailbox slot [seq & 1][my rank]     (NVLink P2P store, system scope)
//                spin on my own slot [seq & 1][t] until its sequence number is seq    (peer t's store landing)
//     block:     max over t -> *value
// Two slots by sequence parity are enough: a rank cannot finish step s+1 before every peer has published s+1, which
// a peer only does after it has consumed step s.  max is exact, so the result is bitwise the one NCCL gives.
// A rank that never shows up trips a clock-based timeout that raises an error flag instead of hanging the GPU.
#include "../../include/exahype_cuda.h"
#include "peer_mail.cuh"

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <vector>

namespace exahype {

struct PeerReducer {
  int world = 0, rank = 0, device = 0;
  PeerMail* mine = nullptr;                 // [2][world], this rank's mailbox (cudaMalloc, IPC-exported)
  std::vector<PeerMail*> peers;             // peer r's mailbox as mapped here (peers[rank] == mine)
  PeerMail** d_peers = nullptr;             // device copy of `peers`
  int* d_error = nullptr;                   // set to 1 by a timed-out wait
  unsigned int* d_ticket = nullptr;         // arrival counter of patch kernels that run the exchange in their epilogue
  unsigned long long seq = 0;
  bool connected = false;
};

namespace {

template <typename T>
__global__ void peer_allreduce_max_kernel(T* value, PeerMail* const* peers, PeerMail* mine, int world, int rank,
                                          unsigned long long seq, long long timeout_cycles, int* error) {
  __shared__ T partial[32];
  const int t = threadIdx.x;
  T v = *value;
  T got = v;
  if (t < world) got = peer_exchange_with<T>(peers, mine, world, rank, t, seq, timeout_cycles, error, v);
  // max over the block (world <= 1024 threads): std::max semantics, NaN-free inputs
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, got, o);
    got = (got < other) ? other : got;
  }
  if ((t & 31) == 0) partial[t >> 5] = got;
  __syncthreads();
  if (t == 0) {
    T m = partial[0];
    for (int w = 1; w < (int)((blockDim.x + 31) >> 5); ++w) m = (m < partial[w]) ? partial[w] : m;
    *value = m;
  }
}

}  // namespace

cudaError_t peer_reducer_create(PeerReducer** out, int world, int rank) {
  PeerReducer* r = new PeerReducer;
  r->world = world; r->rank = rank;
  cudaError_t err = cudaGetDevice(&r->device);
  if (err == cudaSuccess) err = cudaMalloc(&r->mine, sizeof(PeerMail) * 2 * world);
  if (err == cudaSuccess) err = cudaMemset(r->mine, 0, sizeof(PeerMail) * 2 * world);
  if (err == cudaSuccess) err = cudaMalloc(&r->d_peers, sizeof(PeerMail*) * world);
  if (err == cudaSuccess) err = cudaMalloc(&r->d_error, sizeof(int));
  if (err == cudaSuccess) err = cudaMemset(r->d_error, 0, sizeof(int));
  if (err == cudaSuccess) err = cudaMalloc(&r->d_ticket, sizeof(unsigned int));
  if (err == cudaSuccess) err = cudaMemset(r->d_ticket, 0, sizeof(unsigned int));
  if (err != cudaSuccess) { delete r; return err; }
  r->peers.assign(world, nullptr);
  r->peers[rank] = r->mine;
  *out = r;
  return cudaSuccess;
}

cudaError_t peer_reducer_local_handle(PeerReducer* r, void* out64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t err = cudaIpcGetMemHandle(&h, r->mine);
  if (err == cudaSuccess) std::memcpy(out64, &h, sizeof h);
  return err;
}

cudaError_t peer_reducer_connect(PeerReducer* r, const void* all_handles) {
  const char* base = static_cast<const char*>(all_handles);
  for (int p = 0; p < r->world; ++p) {
    if (p == r->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, base + (size_t)p * sizeof h, sizeof h);
    void* mapped = nullptr;
    cudaError_t err = cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) return err;
    r->peers[p] = static_cast<PeerMail*>(mapped);
  }
  cudaError_t err = cudaMemcpy(r->d_peers, r->peers.data(), sizeof(PeerMail*) * r->world, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) r->connected = true;
  return err;
}

cudaError_t peer_reducer_allreduce_max(PeerReducer* r, void* value, int dtype, cudaStream_t stream) {
  if (!r->connected) return cudaErrorNotReady;
  const unsigned long long seq = ++r->seq;
  const int threads = (r->world + 31) / 32 * 32;
  const long long timeout = 20000000000ll;               // ~10 s of SM clocks: a missing rank raises the error flag
  if (dtype == EXAHYPE_DTYPE_F64)
    peer_allreduce_max_kernel<double><<<1, threads, 0, stream>>>(static_cast<double*>(value), r->d_peers, r->mine, r->world,
                                                                 r->rank, seq, timeout, r->d_error);
  else
    peer_allreduce_max_kernel<float><<<1, threads, 0, stream>>>(static_cast<float*>(value), r->d_peers, r->mine, r->world,
                                                                r->rank, seq, timeout, r->d_error);
  return cudaGetLastError();
}

// the arguments of the NEXT exchange, for a patch kernel that runs it in its epilogue (counts as one allreduce_max call)
cudaError_t peer_reducer_next_fused(PeerReducer* r, FvPeerFuse* out) {
  if (!r->connected) return cudaErrorNotReady;
  if (r->world > 32) return cudaErrorNotSupported;      // one lane per peer
  out->peers = r->d_peers;
  out->mine = r->mine;
  out->ticket = r->d_ticket;
  out->error = r->d_error;
  out->seq = ++r->seq;
  out->timeout_cycles = 20000000000ll;
  out->world = r->world;
  out->rank = r->rank;
  return cudaSuccess;
}

cudaError_t peer_reducer_error(PeerReducer* r, int* flag) {
  return cudaMemcpy(flag, r->d_error, sizeof(int), cudaMemcpyDeviceToHost);
}

void peer_reducer_destroy(PeerReducer* r) {
  if (!r) return;
  for (int p = 0; p < r->world; ++p)
    if (p != r->rank && r->peers[p]) cudaIpcCloseMemHandle(r->peers[p]);
  if (r->mine) cudaFree(r->mine);
  if (r->d_peers) cudaFree(r->d_peers);
  if (r->d_error) cudaFree(r->d_error);
  if (r->d_ticket) cudaFree(r->d_ticket);
  delete r;
}

}  // namespace exahype
