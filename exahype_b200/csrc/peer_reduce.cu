// One-shot all-reduce(max) of the admissible-time-step scalar over NVLink peer memory, and the device-resident time
// loop built on it.
//
// The reference has no distributed code; the exchange this path needs is ONE scalar per step, so
// the collective is pure latency.  ncclAllReduce of 8 bytes costs ~17 us per step on 2 GPUs; here every rank owns a
// mailbox [2][world] of 8-byte words in device memory, opened by every peer through CUDA IPC (protocol: peer_mail.cuh).
//
//   PeerReducer  the mailboxes + the blocking exchange as one tiny stream-ordered kernel
//                (peer_allreduce_max_kernel: thread t publishes to peer t and waits for peer t);
//   TimeLoop     dt of step k+1 = cfl_dx / max over all ranks of lambda_max(step k), never leaving the device: the
//                patch kernel of step k publishes its device maximum (its own epilogue where the kernel can, a
//                one-warp kernel behind it otherwise) and the patch kernel of step k+1 consumes the exchange in its
//                prologue.  A step is ONE launch for the warp-per-patch kernel, with no memset and no host round trip.
//
// A rank that never shows up trips a clock-based timeout: the waiting side poisons its result with NaN and raises a
// sticky flag in host-mapped memory, which every later API call on the reducer reports without synchronising.
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
  int* h_error = nullptr;                   // host-mapped: set to 1 by a timed-out wait (sticky)
  int* d_error = nullptr;                   // the same word as the device sees it
  unsigned int* d_ticket = nullptr;         // arrival counter of patch kernels that run the exchange in their epilogue
  unsigned long long seq = 0;               // exchanges started so far
  unsigned long long pending_seq = 0;       // published by a time loop and not consumed yet (0: none)
  long long timeout_cycles = 20000000000ll; // ~10 s of SM clocks
  unsigned long long* d_trace = nullptr;    // [trace_capacity][FV_TRACE_WORDS] globaltimer stamps, optional
  int trace_capacity = 0;
  bool connected = false;
  bool ipc_mapped = false;                  // peers were opened through CUDA IPC (to be closed again)
};

struct TimeLoop {
  PeerReducer* red = nullptr;
  bool owns_reducer = false;
  int dtype = EXAHYPE_DTYPE_F64;
  double cfl_dx = 0.0;
  char* d_dt = nullptr;          // T[2]: d_dt[cur] is the most recently established time step
  int cur = 0;
  void* d_lambda_acc = nullptr;  // T: the patch kernels' atomicMax target, left at zero by whoever publishes it
  char* d_history = nullptr;     // T[capacity][4]: {dt used, global maximum consumed, device maximum} per step
  long long capacity = 0;
  long long steps = 0;
};

namespace {

size_t elem_size(int dtype) { return dtype == EXAHYPE_DTYPE_F64 ? 8 : 4; }

template <typename T>
__global__ void peer_allreduce_max_kernel(T* value, PeerMail* const* peers, PeerMail* mine, int world, int rank,
                                          unsigned long long seq, long long timeout_cycles, int* error) {
  __shared__ T partial[32];
  __shared__ int bad;
  const int t = threadIdx.x;
  if (t == 0) bad = 0;
  __syncthreads();
  T v = *value;
  T got = v;
  if (t < world) got = peer_exchange_with<T>(peers, mine, world, rank, t, seq, timeout_cycles, error, v);
  if (got != got) { bad = 1; got = v; }
  // max over the block (world <= 1024 threads): std::max semantics
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
    *value = bad ? MailBits<T>::poison() : m;       // a peer never arrived: visibly poisoned, not silently rank-local
  }
}

// time loop, one warp: consume the pending exchange (if any) into the loop's dt scalar and the step's record
template <typename T>
__global__ void loop_consume_kernel(const FvPeerFuse pf) {
  peer_loop_dt<T>(pf, (int)threadIdx.x, T(0), true);
}

// time loop, one warp: publish the device maximum the patch kernel(s) of this step accumulated; leaves the accumulator at 0
template <typename T, typename BitsT>
__global__ void loop_publish_kernel(const FvPeerFuse pf, T* lambda_acc) {
  const int lane = (int)threadIdx.x;
  T v = T(0);
  if (lane == 0) {
    peer_trace(pf, pf.seq, FV_TRACE_LAST_WARP);
    v = MailBits<T>::from((unsigned long long)atomicExch(reinterpret_cast<BitsT*>(lambda_acc), (BitsT)0));
  }
  v = __shfl_sync(0xffffffffu, v, 0);
  for (int t = lane; t < pf.world; t += 32) peer_publish_to<T>(pf.peers, pf.world, pf.rank, t, pf.seq, v);
  __syncwarp();
  if (lane == 0) {
    peer_trace(pf, pf.seq, FV_TRACE_PUBLISHED);
    if (pf.record != nullptr) static_cast<T*>(pf.record)[2] = v;
  }
}

void fill_common(const PeerReducer* r, FvPeerFuse* out) {
  out->peers = r->d_peers;
  out->mine = r->mine;
  out->ticket = r->d_ticket;
  out->error = r->d_error;
  out->timeout_cycles = r->timeout_cycles;
  out->world = r->world;
  out->rank = r->rank;
  out->trace = r->d_trace;
  out->trace_capacity = r->trace_capacity;
}

}  // namespace

void peer_reducer_destroy(PeerReducer* r);

cudaError_t peer_reducer_create(PeerReducer** out, int world, int rank) {
  PeerReducer* r = new PeerReducer;
  r->world = world; r->rank = rank;
  r->peers.assign(world, nullptr);
  cudaError_t err = cudaGetDevice(&r->device);
  if (err == cudaSuccess) err = cudaMalloc(&r->mine, sizeof(PeerMail) * 2 * world);
  if (err == cudaSuccess) {   // "exchange 0 / -1": slot 0 zero, slot 1 with the epoch bit set (peer_mail.cuh)
    std::vector<PeerMail> init(2 * (size_t)world, PeerMail{0ull});
    for (int t = 0; t < world; ++t) init[(size_t)world + t].word = kMailEpochBit;
    err = cudaMemcpy(r->mine, init.data(), sizeof(PeerMail) * 2 * world, cudaMemcpyHostToDevice);
  }
  if (err == cudaSuccess) err = cudaMalloc(&r->d_peers, sizeof(PeerMail*) * world);
  if (err == cudaSuccess) err = cudaHostAlloc(&r->h_error, sizeof(int), cudaHostAllocMapped);
  if (err == cudaSuccess) {
    *r->h_error = 0;
    err = cudaHostGetDevicePointer(&r->d_error, r->h_error, 0);
  }
  if (err == cudaSuccess) err = cudaMalloc(&r->d_ticket, sizeof(unsigned int));
  if (err == cudaSuccess) err = cudaMemset(r->d_ticket, 0, sizeof(unsigned int));
  if (err == cudaSuccess) {
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, r->device) == cudaSuccess && khz > 0)
      r->timeout_cycles = 10ll * 1000ll * khz;          // 10 s
  }
  if (err == cudaSuccess && world == 1) {               // a single rank has nobody to connect to
    r->peers[0] = r->mine;
    err = cudaMemcpy(r->d_peers, r->peers.data(), sizeof(PeerMail*), cudaMemcpyHostToDevice);
    r->connected = (err == cudaSuccess);
  }
  if (err != cudaSuccess) { peer_reducer_destroy(r); return err; }
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
    r->ipc_mapped = true;
  }
  cudaError_t err = cudaMemcpy(r->d_peers, r->peers.data(), sizeof(PeerMail*) * r->world, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) r->connected = true;
  return err;
}

// All ranks live in THIS process on one device (tests; several ranks per GPU): the mailboxes are plain device pointers.
cudaError_t peer_reducer_connect_local(PeerReducer* const* all, int world) {
  for (int a = 0; a < world; ++a) {
    PeerReducer* r = all[a];
    if (r == nullptr || r->world != world || r->rank != a) return cudaErrorInvalidValue;
    for (int p = 0; p < world; ++p) r->peers[p] = all[p]->mine;
    cudaError_t err = cudaMemcpy(r->d_peers, r->peers.data(), sizeof(PeerMail*) * world, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) return err;
    r->connected = true;
  }
  return cudaSuccess;
}

int peer_reducer_world(const PeerReducer* r) { return r->world; }
bool peer_reducer_failed(const PeerReducer* r) { return *reinterpret_cast<volatile int*>(r->h_error) != 0; }
bool peer_reducer_pending(const PeerReducer* r) { return r->pending_seq != 0; }

cudaError_t peer_reducer_set_timeout(PeerReducer* r, double seconds) {
  int khz = 0;
  cudaError_t err = cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, r->device);
  if (err != cudaSuccess) return err;
  const double cycles = seconds * 1e3 * khz;
  r->timeout_cycles = cycles < 1.0 ? 1 : (long long)cycles;
  return cudaSuccess;
}

cudaError_t peer_reducer_enable_trace(PeerReducer* r, int capacity) {
  if (r->d_trace) { cudaFree(r->d_trace); r->d_trace = nullptr; r->trace_capacity = 0; }
  if (capacity <= 0) return cudaSuccess;
  const size_t bytes = sizeof(unsigned long long) * FV_TRACE_WORDS * (size_t)capacity;
  cudaError_t err = cudaMalloc(&r->d_trace, bytes);
  if (err == cudaSuccess) err = cudaMemset(r->d_trace, 0, bytes);
  if (err == cudaSuccess) r->trace_capacity = capacity;
  return err;
}

// host copy of the trace rows of exchanges [first_seq, first_seq + count): count * FV_TRACE_WORDS words; synchronises
cudaError_t peer_reducer_read_trace(PeerReducer* r, unsigned long long first_seq, int count, unsigned long long* out) {
  if (!r->d_trace || count > r->trace_capacity) return cudaErrorInvalidValue;
  cudaError_t err = cudaDeviceSynchronize();
  for (int i = 0; i < count && err == cudaSuccess; ++i)
    err = cudaMemcpy(out + (size_t)i * FV_TRACE_WORDS,
                     r->d_trace + ((first_seq + i) % (unsigned long long)r->trace_capacity) * FV_TRACE_WORDS,
                     sizeof(unsigned long long) * FV_TRACE_WORDS, cudaMemcpyDeviceToHost);
  return err;
}

cudaError_t peer_reducer_allreduce_max(PeerReducer* r, void* value, int dtype, cudaStream_t stream) {
  if (!r->connected || r->pending_seq != 0) return cudaErrorNotReady;
  const unsigned long long seq = ++r->seq;
  const int threads = (r->world + 31) / 32 * 32;
  if (dtype == EXAHYPE_DTYPE_F64)
    peer_allreduce_max_kernel<double><<<1, threads, 0, stream>>>(static_cast<double*>(value), r->d_peers, r->mine, r->world,
                                                                 r->rank, seq, r->timeout_cycles, r->d_error);
  else
    peer_allreduce_max_kernel<float><<<1, threads, 0, stream>>>(static_cast<float*>(value), r->d_peers, r->mine, r->world,
                                                                r->rank, seq, r->timeout_cycles, r->d_error);
  return cudaGetLastError();
}

// the arguments of the NEXT exchange, for a patch kernel that runs it (blocking) in its epilogue: counts as one
// allreduce_max call
cudaError_t peer_reducer_next_fused(PeerReducer* r, FvPeerFuse* out) {
  if (!r->connected || r->pending_seq != 0) return cudaErrorNotReady;
  if (r->world > 32) return cudaErrorNotSupported;      // one lane per peer
  fill_common(r, out);
  out->seq = ++r->seq;
  out->mode = FV_PEER_BLOCKING;
  return cudaSuccess;
}

cudaError_t peer_reducer_error(PeerReducer* r, int* flag) {
  *flag = peer_reducer_failed(r) ? 1 : 0;
  return cudaSuccess;
}

void peer_reducer_destroy(PeerReducer* r) {
  if (!r) return;
  if (r->ipc_mapped)
    for (int p = 0; p < r->world; ++p)
      if (p != r->rank && r->peers[p]) cudaIpcCloseMemHandle(r->peers[p]);
  if (r->mine) cudaFree(r->mine);
  if (r->d_peers) cudaFree(r->d_peers);
  if (r->h_error) cudaFreeHost(r->h_error);
  if (r->d_ticket) cudaFree(r->d_ticket);
  if (r->d_trace) cudaFree(r->d_trace);
  delete r;
}

// ------------------------------------------------------------------------------------------------ time loop
void time_loop_destroy(TimeLoop* l) {
  if (!l) return;
  if (l->d_dt) cudaFree(l->d_dt);
  if (l->d_lambda_acc) cudaFree(l->d_lambda_acc);
  if (l->d_history) cudaFree(l->d_history);
  if (l->owns_reducer) peer_reducer_destroy(l->red);
  delete l;
}

cudaError_t time_loop_create(TimeLoop** out, int dtype, PeerReducer* reducer, double cfl_dx, double dt0,
                             long long history_capacity) {
  TimeLoop* l = new TimeLoop;
  l->dtype = dtype; l->cfl_dx = cfl_dx; l->capacity = history_capacity > 0 ? history_capacity : 4096;
  cudaError_t err = cudaSuccess;
  if (reducer) {
    l->red = reducer;
  } else {
    err = peer_reducer_create(&l->red, 1, 0);
    l->owns_reducer = (err == cudaSuccess);
  }
  const size_t es = elem_size(dtype);
  if (err == cudaSuccess) err = cudaMalloc(&l->d_dt, 2 * es);
  if (err == cudaSuccess) err = cudaMalloc(&l->d_lambda_acc, 8);
  if (err == cudaSuccess) err = cudaMemset(l->d_lambda_acc, 0, 8);
  if (err == cudaSuccess) err = cudaMalloc(&l->d_history, (size_t)l->capacity * 4 * es);
  if (err == cudaSuccess) err = cudaMemset(l->d_history, 0, (size_t)l->capacity * 4 * es);
  if (err == cudaSuccess) {
    double d[2] = {dt0, dt0};
    float f[2] = {(float)dt0, (float)dt0};
    err = cudaMemcpy(l->d_dt, dtype == EXAHYPE_DTYPE_F64 ? (const void*)d : (const void*)f, 2 * es, cudaMemcpyHostToDevice);
  }
  if (err != cudaSuccess) { time_loop_destroy(l); return err; }
  *out = l;
  return cudaSuccess;
}

PeerReducer* time_loop_reducer(TimeLoop* l) { return l->red; }
int time_loop_dtype(const TimeLoop* l) { return l->dtype; }
long long time_loop_steps(const TimeLoop* l) { return l->steps; }
void* time_loop_lambda_acc(TimeLoop* l) { return l->d_lambda_acc; }
void* time_loop_dt_device(TimeLoop* l) { return l->d_dt + (size_t)l->cur * elem_size(l->dtype); }

// Arguments of the next step's launch.  in_kernel_publish: the patch kernel publishes in its own epilogue (FV_PEER_LOOP);
// otherwise it only consumes (FV_PEER_CONSUME) and time_loop_publish must follow it.  Advances the loop's state: call
// exactly once per step, right before the launch.
cudaError_t time_loop_next(TimeLoop* l, bool in_kernel_publish, FvPeerFuse* out) {
  PeerReducer* r = l->red;
  if (!r->connected) return cudaErrorNotReady;
  const size_t es = elem_size(l->dtype);
  fill_common(r, out);
  out->mode = in_kernel_publish ? FV_PEER_LOOP : FV_PEER_CONSUME;
  out->consume_seq = r->pending_seq;
  out->dt_in = l->d_dt + (size_t)l->cur * es;
  out->dt_out = l->d_dt + (size_t)(l->cur ^ 1) * es;
  out->record = l->d_history + (size_t)(l->steps % l->capacity) * 4 * es;
  out->cfl_dx = l->cfl_dx;
  out->seq = ++r->seq;                  // the exchange this step publishes
  r->pending_seq = out->seq;
  l->cur ^= 1;
  ++l->steps;
  return cudaSuccess;
}

// one-warp kernels around patch kernels that cannot run their side of the exchange themselves
cudaError_t time_loop_launch_consume(TimeLoop* l, const FvPeerFuse& pf, cudaStream_t stream) {
  if (l->dtype == EXAHYPE_DTYPE_F64) loop_consume_kernel<double><<<1, 32, 0, stream>>>(pf);
  else loop_consume_kernel<float><<<1, 32, 0, stream>>>(pf);
  return cudaGetLastError();
}
cudaError_t time_loop_launch_publish(TimeLoop* l, const FvPeerFuse& pf, cudaStream_t stream) {
  if (l->dtype == EXAHYPE_DTYPE_F64)
    loop_publish_kernel<double, unsigned long long><<<1, 32, 0, stream>>>(pf, static_cast<double*>(l->d_lambda_acc));
  else
    loop_publish_kernel<float, unsigned int><<<1, 32, 0, stream>>>(pf, static_cast<float*>(l->d_lambda_acc));
  return cudaGetLastError();
}

// Consume the pending exchange: afterwards time_loop_dt_device() holds the time step of the NEXT step, and the
// history's tail entry (index `steps`) {that dt, the last step's global maximum}.  No-op when nothing is pending.
cudaError_t time_loop_flush(TimeLoop* l, cudaStream_t stream) {
  PeerReducer* r = l->red;
  if (r->pending_seq == 0) return cudaSuccess;
  const size_t es = elem_size(l->dtype);
  FvPeerFuse pf;
  fill_common(r, &pf);
  pf.mode = FV_PEER_CONSUME;
  pf.consume_seq = r->pending_seq;
  pf.seq = r->pending_seq;
  pf.dt_in = l->d_dt + (size_t)l->cur * es;
  pf.dt_out = l->d_dt + (size_t)(l->cur ^ 1) * es;
  pf.record = l->d_history + (size_t)(l->steps % l->capacity) * 4 * es;
  pf.cfl_dx = l->cfl_dx;
  cudaError_t err = time_loop_launch_consume(l, pf, stream);
  if (err != cudaSuccess) return err;
  r->pending_seq = 0;
  l->cur ^= 1;
  return cudaSuccess;
}

// host copy of history entries [first, first + count): count * 4 values of the loop's dtype; synchronises the device
cudaError_t time_loop_history(TimeLoop* l, long long first, long long count, void* out) {
  if (first < 0 || count < 0 || first + count > l->steps + 1 || count > l->capacity) return cudaErrorInvalidValue;
  const size_t es = elem_size(l->dtype);
  cudaError_t err = cudaDeviceSynchronize();
  for (long long i = 0; i < count && err == cudaSuccess; ++i)
    err = cudaMemcpy(static_cast<char*>(out) + (size_t)i * 4 * es,
                     l->d_history + (size_t)((first + i) % l->capacity) * 4 * es, 4 * es, cudaMemcpyDeviceToHost);
  return err;
}

}  // namespace exahype
