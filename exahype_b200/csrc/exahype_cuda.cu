// libexahype_cuda.so -- C ABI (include/exahype_cuda.h) over the sm_100a kernel instantiations.
//
// Replaces, for a whole batch of patches and on the GPU, the reference's generated
//   void time_step(double* Q, double dt)            ("Unit test/test.h":3, "Unit test/test.cpp":3-111)
// There is deliberately no CPU implementation behind these entry points.
#include "../../include/exahype_cuda.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "fv_registry.h"

namespace exahype {
struct PeerReducer;   // peer_reduce.cu
cudaError_t peer_reducer_create(PeerReducer** out, int world, int rank);
cudaError_t peer_reducer_local_handle(PeerReducer* r, void* out64);
cudaError_t peer_reducer_connect(PeerReducer* r, const void* all_handles);
cudaError_t peer_reducer_allreduce_max(PeerReducer* r, void* value, int dtype, cudaStream_t stream);
cudaError_t peer_reducer_next_fused(PeerReducer* r, FvPeerFuse* out);
cudaError_t peer_reducer_error(PeerReducer* r, int* flag);
cudaError_t peer_reducer_connect_local(PeerReducer* const* all, int world);
cudaError_t peer_reducer_set_timeout(PeerReducer* r, double seconds);
cudaError_t peer_reducer_enable_trace(PeerReducer* r, int capacity);
cudaError_t peer_reducer_read_trace(PeerReducer* r, unsigned long long first_seq, int count, unsigned long long* out);
int peer_reducer_world(const PeerReducer* r);
bool peer_reducer_failed(const PeerReducer* r);
bool peer_reducer_pending(const PeerReducer* r);
void peer_reducer_destroy(PeerReducer* r);
struct TimeLoop;      // peer_reduce.cu
cudaError_t time_loop_create(TimeLoop** out, int dtype, PeerReducer* reducer, double cfl_dx, double dt0, long long history_capacity);
void time_loop_destroy(TimeLoop* l);
PeerReducer* time_loop_reducer(TimeLoop* l);
int time_loop_dtype(const TimeLoop* l);
long long time_loop_steps(const TimeLoop* l);
void* time_loop_lambda_acc(TimeLoop* l);
void* time_loop_dt_device(TimeLoop* l);
cudaError_t time_loop_next(TimeLoop* l, bool in_kernel_publish, FvPeerFuse* out);
cudaError_t time_loop_launch_consume(TimeLoop* l, const FvPeerFuse& pf, cudaStream_t stream);
cudaError_t time_loop_launch_publish(TimeLoop* l, const FvPeerFuse& pf, cudaStream_t stream);
cudaError_t time_loop_flush(TimeLoop* l, cudaStream_t stream);
cudaError_t time_loop_history(TimeLoop* l, long long first, long long count, void* out);
cudaError_t fill_synthetic(const exahype_fv_config* cfg, void* q, long long first_cell, long long n_cells,
                           unsigned long long seed, cudaStream_t stream);   // synthetic.cu
}

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int cuda_fail(cudaError_t err, const char* what) {
  return fail(EXAHYPE_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(err), cudaGetErrorName(err));
}

// a wait of this reducer timed out earlier: everything derived from its exchanges is NaN from then on
int reducer_failed(const exahype::PeerReducer* r) {
  if (r && exahype::peer_reducer_failed(r))
    return fail(EXAHYPE_ERR_TIMEOUT, "a peer never arrived at an all-reduce(max) exchange (wait timed out): the reduced "
                                     "scalar and every time step derived from it are NaN; destroy the reducer");
  return EXAHYPE_OK;
}

const std::vector<exahype::FvEntry>& registry() {
  static const std::vector<exahype::FvEntry> all = [] {
    std::vector<exahype::FvEntry> v;
    for (exahype::FvEntryList l : {exahype::euler3d_entries(), exahype::euler2d_entries(), exahype::swe2d_entries()})
      v.insert(v.end(), l.entries, l.entries + l.count);
    return v;
  }();
  return all;
}

// mirrors KernelBuilder.viable() (reference exahype/KernelBuilder.py:41-48) plus what the stencil needs
int validate(const exahype_fv_config* cfg) {
  if (!cfg) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "cfg is null");
  if (cfg->dim != 2 && cfg->dim != 3) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "check viability of inputs: dim must be 2 or 3 (got %d)", cfg->dim);
  if (cfg->patch_size < 1) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "check viability of inputs: patch_size must be >= 1 (got %d)", cfg->patch_size);
  if (cfg->halo < 0) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "check viability of inputs: halo_size must be >= 0 (got %d)", cfg->halo);
  if (cfg->halo < 1) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "the Rusanov update reads one halo layer: halo_size must be >= 1");
  if (cfg->n_real < 1 || cfg->n_aux < 0) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "n_real must be >= 1 and n_aux >= 0");
  if (cfg->dtype != EXAHYPE_DTYPE_F64 && cfg->dtype != EXAHYPE_DTYPE_F32) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "unknown dtype %d", cfg->dtype);
  if (cfg->model != EXAHYPE_MODEL_EULER && cfg->model != EXAHYPE_MODEL_SWE && cfg->model != EXAHYPE_MODEL_SWE_SOURCE)
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "unknown model %d", cfg->model);
  if (cfg->flags & ~(EXAHYPE_FLAG_DISSIPATION_ALL | EXAHYPE_FLAG_OUTPUT_UNHALOED | EXAHYPE_FLAG_LAMBDA_ACCUMULATE |
                     EXAHYPE_FLAG_KERNEL_CELL | EXAHYPE_FLAG_FAST_ARITHMETIC | EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY))
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "unknown flag bits 0x%x", cfg->flags);
  if ((cfg->flags & EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY) && !(cfg->flags & EXAHYPE_FLAG_OUTPUT_UNHALOED))
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY is a form of the un-haloed output: set EXAHYPE_FLAG_OUTPUT_UNHALOED too");
  return EXAHYPE_OK;
}

bool same_shape(const exahype_fv_config& c, const exahype_fv_config* cfg) {
  return c.model == cfg->model && c.dtype == cfg->dtype && c.dim == cfg->dim && c.patch_size == cfg->patch_size &&
         c.halo == cfg->halo && c.n_real == cfg->n_real && c.n_aux == cfg->n_aux;
}

const exahype::FvEntry* find(const exahype_fv_config* cfg) {
  if (cfg->flags & EXAHYPE_FLAG_FAST_ARITHMETIC) {      // a permission, not a demand: shapes without a fast build keep
    const exahype::FvEntryList fast = exahype::fast_entries();   // the reference arithmetic
    for (int i = 0; i < fast.count; ++i)
      if (same_shape(fast.entries[i].cfg, cfg) && !(cfg->flags & EXAHYPE_FLAG_KERNEL_CELL)) return &fast.entries[i];
  }
  for (const exahype::FvEntry& e : registry())
    if (same_shape(e.cfg, cfg)) return &e;
  return nullptr;
}

int variant_of(unsigned flags) {
  return ((flags & EXAHYPE_FLAG_DISSIPATION_ALL) ? 1 : 0) | ((flags & EXAHYPE_FLAG_OUTPUT_UNHALOED) ? 2 : 0);
}

// The kernel a call runs: the shape's default one, its thread-per-cell alternative (EXAHYPE_FLAG_KERNEL_CELL), or the
// form whose un-haloed output carries the unknowns only.  Shapes without auxiliary variables have one un-haloed layout.
struct Picked {
  exahype::FvLaunchFn launch;
  exahype::FvPrepareFn prepare;
};
int pick(const exahype::FvEntry* e, const exahype_fv_config* cfg, Picked* out) {
  const int var = variant_of(cfg->flags);
  if ((cfg->flags & EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY) && cfg->n_aux > 0) {
    const int da = var & 1;
    if ((cfg->flags & EXAHYPE_FLAG_KERNEL_CELL) || !e->unknowns_launch[da])
      return fail(EXAHYPE_ERR_NO_INSTANTIATION, "EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY: this shape's kernel has no such output form "
                                                "(committed for the row-marching kernel of the shallow-water families)");
    *out = {e->unknowns_launch[da], e->unknowns_prepare[da]};
    return EXAHYPE_OK;
  }
  const bool alt = (cfg->flags & EXAHYPE_FLAG_KERNEL_CELL) && e->alt_launch[var];
  *out = {alt ? e->alt_launch[var] : e->launch[var], alt ? e->alt_prepare[var] : e->prepare[var]};
  return EXAHYPE_OK;
}

size_t elem_size(int dtype) { return dtype == EXAHYPE_DTYPE_F64 ? 8 : 4; }

long long ipow_ll(long long b, int e) {
  long long r = 1;
  while (e-- > 0) r *= b;
  return r;
}

int lookup(const exahype_fv_config* cfg, const exahype::FvEntry** entry) {
  int rc = validate(cfg);
  if (rc) return rc;
  *entry = find(cfg);
  if (!*entry)
    return fail(EXAHYPE_ERR_NO_INSTANTIATION,
                "no committed instantiation for model=%d dtype=%d dim=%d patch_size=%d halo=%d n_real=%d n_aux=%d; "
                "generate one with exahype.printers.CUDAPrinter",
                cfg->model, cfg->dtype, cfg->dim, cfg->patch_size, cfg->halo, cfg->n_real, cfg->n_aux);
  return EXAHYPE_OK;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int exahype_cuda_version(void) { return EXAHYPE_CUDA_ABI_VERSION; }

const char* exahype_cuda_last_error(void) { return g_last_error.c_str(); }

int exahype_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int exahype_cuda_fv_supported(const exahype_fv_config* cfg) { return (cfg && find(cfg)) ? 1 : 0; }

int exahype_cuda_fv_list(exahype_fv_config* out, int capacity) {
  const auto& r = registry();
  for (int i = 0; i < (int)r.size() && i < capacity && out; ++i) out[i] = r[i].cfg;
  return (int)r.size();
}

int64_t exahype_cuda_launch_count(void) { return g_launches.load(); }

int exahype_cuda_fill_synthetic(const exahype_fv_config* cfg, void* q, int64_t first_cell, int64_t n_cells,
                                uint64_t seed, void* stream) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (n_cells < 0 || first_cell < 0) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "first_cell / n_cells must be >= 0");
  if (n_cells == 0) return EXAHYPE_OK;
  if (!q) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "q must not be null");
  const int nv = cfg->n_real + cfg->n_aux;
  if ((cfg->model == EXAHYPE_MODEL_EULER && cfg->n_real < cfg->dim + 2) || (cfg->model != EXAHYPE_MODEL_EULER && nv < 3))
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "too few variables for the model's synthetic state");
  cudaError_t err = exahype::fill_synthetic(cfg, q, first_cell, n_cells, seed, static_cast<cudaStream_t>(stream));
  if (err != cudaSuccess) return cuda_fail(err, "fill_synthetic_kernel launch");
  g_launches.fetch_add(1);
  return EXAHYPE_OK;
}

int exahype_cuda_fv_launch_info(const exahype_fv_config* cfg, int64_t n_patches, int* grid, int* block,
                                int* smem_bytes, int* patches_per_tile) {
  const exahype::FvEntry* e = nullptr;
  int rc = lookup(cfg, &e);
  if (rc) return rc;
  exahype::FvLaunchInfo info;
  Picked k;
  if ((rc = pick(e, cfg, &k))) return rc;
  cudaError_t err = k.prepare(&info, n_patches);
  if (err != cudaSuccess) return cuda_fail(err, "exahype_cuda_fv_launch_info");
  if (grid) *grid = info.grid;
  if (block) *block = info.block;
  if (smem_bytes) *smem_bytes = info.smem_bytes;
  if (patches_per_tile) *patches_per_tile = info.patches_per_tile;
  return EXAHYPE_OK;
}

int exahype_cuda_fv_step(const exahype_fv_config* cfg, const void* q_in, void* q_out, int64_t n_patches, double dt,
                         void* lambda_patch, void* lambda_max, void* stream) {
  const exahype::FvEntry* e = nullptr;
  int rc = lookup(cfg, &e);
  if (rc) return rc;
  if (n_patches < 0) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "n_patches must be >= 0 (got %lld)", (long long)n_patches);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (lambda_max && !(cfg->flags & EXAHYPE_FLAG_LAMBDA_ACCUMULATE)) {
    cudaError_t err = cudaMemsetAsync(lambda_max, 0, elem_size(cfg->dtype), s);
    if (err != cudaSuccess) return cuda_fail(err, "cudaMemsetAsync(lambda_max)");
  }
  if (n_patches == 0) return EXAHYPE_OK;
  if (!q_in || !q_out) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "q_in / q_out must not be null");
  if ((reinterpret_cast<uintptr_t>(q_in) & 15) || (reinterpret_cast<uintptr_t>(q_out) & 15))
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "q_in / q_out must be 16-byte aligned (TMA bulk copies)");
  if ((cfg->flags & EXAHYPE_FLAG_OUTPUT_UNHALOED) && q_in == q_out)
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "un-haloed output cannot alias the haloed input");
  Picked k;
  if ((rc = pick(e, cfg, &k))) return rc;
  cudaError_t err = k.launch(q_in, q_out, n_patches, dt, lambda_patch, lambda_max, s, nullptr);
  if (err != cudaSuccess) return cuda_fail(err, "fv_step_kernel launch");
  g_launches.fetch_add(1);
  return EXAHYPE_OK;
}

int exahype_cuda_fv_step_allreduce(const exahype_fv_config* cfg, void* reducer, const void* q_in, void* q_out,
                                   int64_t n_patches, double dt, void* lambda_patch, void* lambda_max, void* stream) {
  const exahype::FvEntry* e = nullptr;
  int rc = lookup(cfg, &e);
  if (rc) return rc;
  if (!reducer || !lambda_max) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "reducer / lambda_max must not be null");
  Picked k;
  if ((rc = pick(e, cfg, &k))) return rc;
  exahype::FvLaunchInfo info;
  cudaError_t err = k.prepare(&info, n_patches > 0 ? n_patches : 1);
  if (err != cudaSuccess) return cuda_fail(err, "exahype_cuda_fv_step_allreduce (launch info)");
  exahype::PeerReducer* r = static_cast<exahype::PeerReducer*>(reducer);
  if ((rc = reducer_failed(r))) return rc;
  if (!info.fused_allreduce || n_patches <= 0 || exahype::peer_reducer_world(r) > 32) {
    // kernels without the fused epilogue, empty shards, more than 32 ranks (the epilogue has one lane per peer): the
    // step, then the stand-alone one-shot kernel
    rc = exahype_cuda_fv_step(cfg, q_in, q_out, n_patches, dt, lambda_patch, lambda_max, stream);
    if (rc) return rc;
    return exahype_cuda_peer_reducer_allreduce_max(reducer, lambda_max, cfg->dtype, stream);
  }
  if (!q_in || !q_out) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "q_in / q_out must not be null");
  if ((reinterpret_cast<uintptr_t>(q_in) & 15) || (reinterpret_cast<uintptr_t>(q_out) & 15))
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "q_in / q_out must be 16-byte aligned (TMA bulk copies)");
  if ((cfg->flags & EXAHYPE_FLAG_OUTPUT_UNHALOED) && q_in == q_out)
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "un-haloed output cannot alias the haloed input");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!(cfg->flags & EXAHYPE_FLAG_LAMBDA_ACCUMULATE)) {
    err = cudaMemsetAsync(lambda_max, 0, elem_size(cfg->dtype), s);
    if (err != cudaSuccess) return cuda_fail(err, "cudaMemsetAsync(lambda_max)");
  }
  exahype::FvGatherRaw g = {nullptr, nullptr, nullptr, {}, nullptr, nullptr, nullptr};
  err = exahype::peer_reducer_next_fused(r, &g.peer);
  if (err != cudaSuccess) return cuda_fail(err, "peer reducer not connected, or a time loop's exchange is still pending (flush it)");
  err = k.launch(q_in, q_out, n_patches, dt, lambda_patch, lambda_max, s, &g);
  if (err != cudaSuccess) return cuda_fail(err, "fv_step_kernel launch (fused all-reduce)");
  g_launches.fetch_add(1);
  return EXAHYPE_OK;
}

int exahype_cuda_fv_step_cell_data(const exahype_fv_config* cfg, const exahype_cell_data* cells, double dt,
                                   void* lambda_max, void* stream) {
  const exahype::FvEntry* e = nullptr;
  int rc = lookup(cfg, &e);
  if (rc) return rc;
  if (!cells) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "cells is null");
  if ((cfg->flags & EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY) && cfg->n_aux > 0)
    return fail(EXAHYPE_ERR_NO_INSTANTIATION, "EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY: CellData::QOut carries the auxiliary variables");
  if (cells->n_patches < 0) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "n_patches must be >= 0 (got %lld)", (long long)cells->n_patches);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (lambda_max && !(cfg->flags & EXAHYPE_FLAG_LAMBDA_ACCUMULATE)) {
    cudaError_t err = cudaMemsetAsync(lambda_max, 0, elem_size(cfg->dtype), s);
    if (err != cudaSuccess) return cuda_fail(err, "cudaMemsetAsync(lambda_max)");
  }
  if (cells->n_patches == 0) return EXAHYPE_OK;
  if (!cells->q_in || !cells->q_out) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "cells->q_in / cells->q_out must not be null");
  const exahype::FvGatherRaw g = {cells->q_in, cells->q_out, cells->dt, {}, cells->cell_centre, cells->cell_size, cells->t};
  const int var = variant_of(cfg->flags);   // the CellData form exists for the shape's default kernel
  cudaError_t err = e->gather_launch[var](nullptr, nullptr, cells->n_patches, dt, cells->max_eigenvalue, lambda_max, s, &g);
  if (err != cudaSuccess) return cuda_fail(err, "fv_step_kernel launch (cell data)");
  g_launches.fetch_add(1);
  return EXAHYPE_OK;
}

#define EXAHYPE_NAMED_ENTRY(NAME, MODEL, DTYPE, DIM, NR, T)                                                     \
  int NAME(const T* q_in, T* q_out, int64_t n_patches, int patch_size, int halo, int n_aux, T dt,               \
           T* lambda_patch, T* lambda_max, unsigned flags, void* stream) {                                      \
    exahype_fv_config cfg = {MODEL, DTYPE, DIM, patch_size, halo, NR, n_aux, flags};                            \
    return exahype_cuda_fv_step(&cfg, q_in, q_out, n_patches, (double)dt, lambda_patch, lambda_max, stream);    \
  }
EXAHYPE_NAMED_ENTRY(exahype_cuda_fv_step_euler_2d_f64, EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, 2, 4, double)
EXAHYPE_NAMED_ENTRY(exahype_cuda_fv_step_euler_3d_f64, EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, 3, 5, double)
EXAHYPE_NAMED_ENTRY(exahype_cuda_fv_step_swe_2d_f64, EXAHYPE_MODEL_SWE, EXAHYPE_DTYPE_F64, 2, 3, double)
EXAHYPE_NAMED_ENTRY(exahype_cuda_fv_step_swe_2d_f32, EXAHYPE_MODEL_SWE, EXAHYPE_DTYPE_F32, 2, 3, float)

// ------------------------------------------------------------------------------------------------
// time_step on HOST memory: chunked H2D -> kernel -> D2H pipeline over `depth` stream slots
namespace {

struct HostPipeline {
  int device = -1;
  int depth = 0;
  size_t in_capacity = 0, out_capacity = 0, lam_capacity = 0;
  std::vector<cudaStream_t> streams;
  std::vector<void*> d_in, d_out, d_lam;
  void* d_lam_max = nullptr;
  void release() {
    for (void* p : d_in) cudaFree(p);
    for (void* p : d_out) cudaFree(p);
    for (void* p : d_lam) cudaFree(p);
    for (cudaStream_t s : streams) cudaStreamDestroy(s);
    if (d_lam_max) cudaFree(d_lam_max);
    d_in.clear(); d_out.clear(); d_lam.clear(); streams.clear();
    d_lam_max = nullptr; depth = 0; in_capacity = out_capacity = lam_capacity = 0; device = -1;
  }
};

// One pipeline per device, each behind its own mutex: host threads driving different GPUs neither serialise nor evict
// each other's staging buffers (two calls for the SAME device do serialise: they share its staging ring).
struct DevicePipeline {
  std::mutex mutex;
  HostPipeline pipe;
};
constexpr int kMaxDevices = 64;
DevicePipeline g_pipes[kMaxDevices];
std::atomic<long long> g_chunk_patches{0};   // 0: derive from a ~32 MiB input chunk
std::atomic<int> g_depth{3};

}  // namespace

int exahype_cuda_host_pipeline_configure(int64_t chunk_patches, int depth) {
  if (chunk_patches < 0 || depth < 0 || depth > 8) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "bad pipeline configuration");
  g_chunk_patches = chunk_patches;
  g_depth = depth > 0 ? depth : 3;
  return EXAHYPE_OK;
}

int exahype_cuda_host_pipeline_release(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return EXAHYPE_OK; }   // no device: nothing was cached
  if (dev < 0 || dev >= kMaxDevices) return EXAHYPE_OK;
  std::lock_guard<std::mutex> lock(g_pipes[dev].mutex);
  g_pipes[dev].pipe.release();
  return EXAHYPE_OK;
}

int exahype_cuda_time_step_host(const exahype_fv_config* cfg, const void* q_host, void* q_out_host, int64_t n_patches,
                                double dt, void* lambda_patch_host, void* lambda_max_host) {
  const exahype::FvEntry* e = nullptr;
  int rc = lookup(cfg, &e);
  if (rc) return rc;
  if (n_patches < 0) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "n_patches must be >= 0");
  if (n_patches > 0 && (!q_host || !q_out_host)) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "host buffers must not be null");
  const bool unhaloed = cfg->flags & EXAHYPE_FLAG_OUTPUT_UNHALOED;
  if (unhaloed && q_host == q_out_host) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "un-haloed output cannot alias the haloed input");

  const size_t es = elem_size(cfg->dtype);
  const int nv = cfg->n_real + cfg->n_aux;
  const size_t in_patch = (size_t)ipow_ll(cfg->patch_size + 2 * cfg->halo, cfg->dim) * nv * es;
  const int out_nv = (cfg->flags & EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY) ? cfg->n_real : nv;
  const size_t out_patch = unhaloed ? (size_t)ipow_ll(cfg->patch_size, cfg->dim) * out_nv * es : in_patch;

  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return cuda_fail(err, "cudaGetDevice");
  if (dev < 0 || dev >= kMaxDevices) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_pipes[dev].mutex);
  const long long configured = g_chunk_patches.load();
  long long chunk = configured > 0 ? configured : std::max<long long>(1, (long long)((32u << 20) / in_patch));
  chunk = std::min<long long>(chunk, std::max<long long>(n_patches, 1));
  const int depth = g_depth.load();
  HostPipeline& p = g_pipes[dev].pipe;
  if (p.device != dev || p.depth != depth || p.in_capacity < (size_t)chunk * in_patch ||
      p.out_capacity < (size_t)chunk * out_patch || p.lam_capacity < (size_t)chunk * es) {
    p.release();
    p.device = dev; p.depth = depth;
    p.in_capacity = (size_t)chunk * in_patch;
    p.out_capacity = (size_t)chunk * std::max(in_patch, out_patch);
    p.lam_capacity = (size_t)chunk * es;
    for (int i = 0; i < depth; ++i) {
      cudaStream_t s; void *a = nullptr, *b = nullptr, *c = nullptr;
      if ((err = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess) { p.release(); return cuda_fail(err, "cudaStreamCreate"); }
      p.streams.push_back(s);
      if ((err = cudaMalloc(&a, p.in_capacity)) != cudaSuccess) { p.release(); return cuda_fail(err, "cudaMalloc(staging in)"); }
      p.d_in.push_back(a);
      if ((err = cudaMalloc(&b, p.out_capacity)) != cudaSuccess) { p.release(); return cuda_fail(err, "cudaMalloc(staging out)"); }
      p.d_out.push_back(b);
      if ((err = cudaMalloc(&c, p.lam_capacity)) != cudaSuccess) { p.release(); return cuda_fail(err, "cudaMalloc(lambda)"); }
      p.d_lam.push_back(c);
    }
    if ((err = cudaMalloc(&p.d_lam_max, 8)) != cudaSuccess) { p.release(); return cuda_fail(err, "cudaMalloc(lambda_max)"); }
  }

  if ((err = cudaMemsetAsync(p.d_lam_max, 0, 8, p.streams[0])) != cudaSuccess) return cuda_fail(err, "cudaMemsetAsync");
  if ((err = cudaStreamSynchronize(p.streams[0])) != cudaSuccess) return cuda_fail(err, "cudaStreamSynchronize");

  exahype_fv_config dev_cfg = *cfg;
  dev_cfg.flags |= EXAHYPE_FLAG_LAMBDA_ACCUMULATE;
  const char* src = static_cast<const char*>(q_host);
  char* dst = static_cast<char*>(q_out_host);
  long long slot_idx = 0;
  for (long long first = 0; first < n_patches; first += chunk, ++slot_idx) {
    const long long n = std::min<long long>(chunk, n_patches - first);
    const int k = (int)(slot_idx % depth);
    cudaStream_t s = p.streams[k];
    // stream order makes reuse of slot k safe: its previous D2H is ahead of this H2D on the same stream
    if ((err = cudaMemcpyAsync(p.d_in[k], src + (size_t)first * in_patch, (size_t)n * in_patch, cudaMemcpyHostToDevice, s)) != cudaSuccess)
      return cuda_fail(err, "cudaMemcpyAsync(H2D)");
    // haloed output: update in place in the staging buffer and copy whole patches back (halos are unchanged)
    void* d_out = unhaloed ? p.d_out[k] : p.d_in[k];
    rc = exahype_cuda_fv_step(&dev_cfg, p.d_in[k], d_out, n, dt, lambda_patch_host ? p.d_lam[k] : nullptr,
                              p.d_lam_max, s);
    if (rc) return rc;
    if ((err = cudaMemcpyAsync(dst + (size_t)first * out_patch, d_out, (size_t)n * out_patch, cudaMemcpyDeviceToHost, s)) != cudaSuccess)
      return cuda_fail(err, "cudaMemcpyAsync(D2H)");
    if (lambda_patch_host &&
        (err = cudaMemcpyAsync(static_cast<char*>(lambda_patch_host) + (size_t)first * es, p.d_lam[k], (size_t)n * es,
                               cudaMemcpyDeviceToHost, s)) != cudaSuccess)
      return cuda_fail(err, "cudaMemcpyAsync(lambda D2H)");
  }
  for (cudaStream_t s : p.streams)
    if ((err = cudaStreamSynchronize(s)) != cudaSuccess) return cuda_fail(err, "cudaStreamSynchronize");
  if (lambda_max_host) {
    if ((err = cudaMemcpy(lambda_max_host, p.d_lam_max, es, cudaMemcpyDeviceToHost)) != cudaSuccess)
      return cuda_fail(err, "cudaMemcpy(lambda_max)");
  }
  return EXAHYPE_OK;
}

// ------------------------------------------------------------------------------------------------
// NCCL all-reduce(max) of the admissible-time-step scalar.  libnccl is loaded lazily so the library itself has no
// link-time dependency on it (CPU-only hosts can still load libexahype_cuda.so and list its symbols).
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclFloat32_ = 7, ncclFloat64_ = 8 };   // nccl.h: ncclFloat = 7, ncclDouble = 8
enum { ncclMax_ = 2 };                         // nccl.h: ncclSum 0, ncclProd 1, ncclMax 2, ncclMin 3

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string why;
};

NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) {
      a.why = std::string("cannot load libnccl: ") + dlerror();
      return a;
    }
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.handle, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.handle, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.handle, "ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(a.handle, "ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.handle, "ncclGetErrorString"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString) {
      a.why = "libnccl is missing a required symbol";
      a.handle = nullptr;
    }
    return a;
  }();
  return api;
}

int nccl_fail(int code, const char* what) {
  return fail(EXAHYPE_ERR_NCCL, "%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(code) : "?");
}

}  // namespace

int exahype_cuda_nccl_unique_id(void* id128) {
  if (!id128) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "id128 is null");
  if (!nccl().handle) return fail(EXAHYPE_ERR_UNAVAILABLE, "%s", nccl().why.c_str());
  ncclUniqueId id;
  int rc = nccl().GetUniqueId(&id);
  if (rc != ncclSuccess_) return nccl_fail(rc, "ncclGetUniqueId");
  std::memcpy(id128, &id, sizeof id);
  return EXAHYPE_OK;
}

int exahype_cuda_comm_init(void** comm, const void* id128, int world_size, int rank) {
  if (!comm || !id128 || world_size < 1 || rank < 0 || rank >= world_size)
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "bad communicator arguments (world_size=%d rank=%d)", world_size, rank);
  if (!nccl().handle) return fail(EXAHYPE_ERR_UNAVAILABLE, "%s", nccl().why.c_str());
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  ncclComm_t c = nullptr;
  int rc = nccl().CommInitRank(&c, world_size, id, rank);
  if (rc != ncclSuccess_) return nccl_fail(rc, "ncclCommInitRank");
  *comm = c;
  return EXAHYPE_OK;
}

int exahype_cuda_comm_destroy(void* comm) {
  if (!comm) return EXAHYPE_OK;
  if (!nccl().handle) return fail(EXAHYPE_ERR_UNAVAILABLE, "%s", nccl().why.c_str());
  int rc = nccl().CommDestroy(static_cast<ncclComm_t>(comm));
  if (rc != ncclSuccess_) return nccl_fail(rc, "ncclCommDestroy");
  return EXAHYPE_OK;
}

int exahype_cuda_allreduce_max(void* comm, void* values, int64_t count, int dtype, void* stream) {
  if (!comm || !values || count < 1) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "bad all-reduce arguments");
  if (dtype != EXAHYPE_DTYPE_F64 && dtype != EXAHYPE_DTYPE_F32) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
  if (!nccl().handle) return fail(EXAHYPE_ERR_UNAVAILABLE, "%s", nccl().why.c_str());
  int rc = nccl().AllReduce(values, values, (size_t)count, dtype == EXAHYPE_DTYPE_F64 ? ncclFloat64_ : ncclFloat32_,
                            ncclMax_, static_cast<ncclComm_t>(comm), static_cast<cudaStream_t>(stream));
  if (rc != ncclSuccess_) return nccl_fail(rc, "ncclAllReduce(max)");
  return EXAHYPE_OK;
}

// ------------------------------------------------------------------------------------------------
// one-shot all-reduce(max) over NVLink peer memory (peer_reduce.cu)
int exahype_cuda_peer_reducer_create(void** reducer, int world_size, int rank) {
  if (!reducer || world_size < 1 || world_size > 1024 || rank < 0 || rank >= world_size)
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "bad reducer arguments (world_size=%d rank=%d)", world_size, rank);
  exahype::PeerReducer* r = nullptr;
  cudaError_t err = exahype::peer_reducer_create(&r, world_size, rank);
  if (err != cudaSuccess) return cuda_fail(err, "peer_reducer_create");
  *reducer = r;
  return EXAHYPE_OK;
}

int exahype_cuda_peer_reducer_local_handle(void* reducer, void* handle64) {
  if (!reducer || !handle64) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "null reducer / handle");
  cudaError_t err = exahype::peer_reducer_local_handle(static_cast<exahype::PeerReducer*>(reducer), handle64);
  if (err != cudaSuccess) return cuda_fail(err, "cudaIpcGetMemHandle");
  return EXAHYPE_OK;
}

int exahype_cuda_peer_reducer_connect(void* reducer, const void* all_handles) {
  if (!reducer || !all_handles) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "null reducer / handles");
  cudaError_t err = exahype::peer_reducer_connect(static_cast<exahype::PeerReducer*>(reducer), all_handles);
  if (err != cudaSuccess) return cuda_fail(err, "cudaIpcOpenMemHandle (peer mailbox)");
  return EXAHYPE_OK;
}

int exahype_cuda_peer_reducer_allreduce_max(void* reducer, void* value, int dtype, void* stream) {
  if (!reducer || !value) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "null reducer / value");
  if (dtype != EXAHYPE_DTYPE_F64 && dtype != EXAHYPE_DTYPE_F32) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
  if (int rc = reducer_failed(static_cast<exahype::PeerReducer*>(reducer))) return rc;
  cudaError_t err = exahype::peer_reducer_allreduce_max(static_cast<exahype::PeerReducer*>(reducer), value, dtype,
                                                        static_cast<cudaStream_t>(stream));
  if (err == cudaErrorNotReady)
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "peer reducer not connected, or a time loop's exchange is still pending (flush it)");
  if (err != cudaSuccess) return cuda_fail(err, "peer_allreduce_max_kernel launch");
  g_launches.fetch_add(1);
  return EXAHYPE_OK;
}

int exahype_cuda_peer_reducer_status(void* reducer, int* flag) {
  if (!reducer || !flag) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "null reducer / flag");
  cudaError_t err = exahype::peer_reducer_error(static_cast<exahype::PeerReducer*>(reducer), flag);
  if (err != cudaSuccess) return cuda_fail(err, "peer reducer status");
  return EXAHYPE_OK;
}

int exahype_cuda_peer_reducer_set_timeout(void* reducer, double seconds) {
  if (!reducer || !(seconds > 0.0)) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "null reducer / non-positive timeout");
  cudaError_t err = exahype::peer_reducer_set_timeout(static_cast<exahype::PeerReducer*>(reducer), seconds);
  if (err != cudaSuccess) return cuda_fail(err, "peer reducer timeout");
  return EXAHYPE_OK;
}

int exahype_cuda_peer_reducer_connect_local(void* const* reducers, int world_size) {
  if (!reducers || world_size < 1) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "null reducers / bad world size");
  cudaError_t err = exahype::peer_reducer_connect_local(reinterpret_cast<exahype::PeerReducer* const*>(reducers), world_size);
  if (err != cudaSuccess) return cuda_fail(err, "peer_reducer_connect_local (reducers must be ranks 0..world-1 of one world)");
  return EXAHYPE_OK;
}

int exahype_cuda_peer_reducer_enable_trace(void* reducer, int capacity) {
  if (!reducer || capacity < 0) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "null reducer / negative capacity");
  cudaError_t err = exahype::peer_reducer_enable_trace(static_cast<exahype::PeerReducer*>(reducer), capacity);
  if (err != cudaSuccess) return cuda_fail(err, "peer reducer trace buffer");
  return EXAHYPE_OK;
}

int exahype_cuda_peer_reducer_read_trace(void* reducer, uint64_t first_seq, int count, uint64_t* out) {
  if (!reducer || !out || count < 0) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "null reducer / out");
  static_assert(sizeof(uint64_t) == sizeof(unsigned long long), "uint64_t");
  cudaError_t err = exahype::peer_reducer_read_trace(static_cast<exahype::PeerReducer*>(reducer), first_seq, count,
                                                     reinterpret_cast<unsigned long long*>(out));
  if (err != cudaSuccess) return cuda_fail(err, "peer reducer trace (enabled? count <= capacity?)");
  return EXAHYPE_OK;
}

int exahype_cuda_peer_reducer_destroy(void* reducer) {
  exahype::peer_reducer_destroy(static_cast<exahype::PeerReducer*>(reducer));
  return EXAHYPE_OK;
}

// ------------------------------------------------------------------------------------------------
// device-resident time loop (peer_reduce.cu, peer_mail.cuh)
int exahype_cuda_time_loop_create(void** loop, int dtype, void* reducer, double cfl_dx, double dt0,
                                  int64_t history_capacity) {
  if (!loop) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "loop is null");
  if (dtype != EXAHYPE_DTYPE_F64 && dtype != EXAHYPE_DTYPE_F32) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
  if (!(cfl_dx > 0.0) || history_capacity < 0) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "cfl_dx must be > 0 and history_capacity >= 0");
  if (int rc = reducer_failed(static_cast<exahype::PeerReducer*>(reducer))) return rc;
  if (reducer && exahype::peer_reducer_pending(static_cast<exahype::PeerReducer*>(reducer)))
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "the reducer still carries another time loop's exchange: flush that loop first");
  exahype::TimeLoop* l = nullptr;
  cudaError_t err = exahype::time_loop_create(&l, dtype, static_cast<exahype::PeerReducer*>(reducer), cfl_dx, dt0, history_capacity);
  if (err != cudaSuccess) return cuda_fail(err, "time_loop_create");
  *loop = l;
  return EXAHYPE_OK;
}

int exahype_cuda_fv_step_time_loop(const exahype_fv_config* cfg, void* loop, const void* q_in, void* q_out,
                                   int64_t n_patches, void* lambda_patch, void* stream) {
  const exahype::FvEntry* e = nullptr;
  int rc = lookup(cfg, &e);
  if (rc) return rc;
  if (!loop) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "loop is null");
  exahype::TimeLoop* l = static_cast<exahype::TimeLoop*>(loop);
  exahype::PeerReducer* r = exahype::time_loop_reducer(l);
  if ((rc = reducer_failed(r))) return rc;
  if (cfg->dtype != exahype::time_loop_dtype(l)) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "the loop's dtype differs from cfg->dtype");
  if (cfg->flags & EXAHYPE_FLAG_LAMBDA_ACCUMULATE) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "the loop owns lambda_max: no LAMBDA_ACCUMULATE");
  if (n_patches < 0) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "n_patches must be >= 0 (got %lld)", (long long)n_patches);
  if (n_patches > 0) {
    if (!q_in || !q_out) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "q_in / q_out must not be null");
    if ((reinterpret_cast<uintptr_t>(q_in) & 15) || (reinterpret_cast<uintptr_t>(q_out) & 15))
      return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "q_in / q_out must be 16-byte aligned (TMA bulk copies)");
    if ((cfg->flags & EXAHYPE_FLAG_OUTPUT_UNHALOED) && q_in == q_out)
      return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "un-haloed output cannot alias the haloed input");
  }
  Picked k;
  if ((rc = pick(e, cfg, &k))) return rc;
  exahype::FvLaunchInfo info;
  cudaError_t err = k.prepare(&info, n_patches > 0 ? n_patches : 1);
  if (err != cudaSuccess) return cuda_fail(err, "exahype_cuda_fv_step_time_loop (launch info)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool in_kernel = info.fused_allreduce && n_patches > 0 && exahype::peer_reducer_world(r) <= 32;
  exahype::FvGatherRaw g = {nullptr, nullptr, nullptr, {}, nullptr, nullptr, nullptr};
  err = exahype::time_loop_next(l, in_kernel, &g.peer);
  if (err != cudaSuccess) return cuda_fail(err, "time loop: reducer not connected");
  if (n_patches > 0) {
    // dt by value is unused (the kernel reads / derives it on the device); lambda_max is the loop's accumulator
    err = k.launch(q_in, q_out, n_patches, 0.0, lambda_patch, exahype::time_loop_lambda_acc(l), s, &g);
    if (err != cudaSuccess) return cuda_fail(err, "fv_step_kernel launch (time loop)");
    g_launches.fetch_add(1);
  } else {
    err = exahype::time_loop_launch_consume(l, g.peer, s);       // an empty shard still takes part in the exchange
    if (err != cudaSuccess) return cuda_fail(err, "loop_consume_kernel launch");
    g_launches.fetch_add(1);
  }
  if (!in_kernel) {
    err = exahype::time_loop_launch_publish(l, g.peer, s);
    if (err != cudaSuccess) return cuda_fail(err, "loop_publish_kernel launch");
    g_launches.fetch_add(1);
  }
  return EXAHYPE_OK;
}

int exahype_cuda_time_loop_flush(void* loop, void* stream) {
  if (!loop) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "loop is null");
  exahype::TimeLoop* l = static_cast<exahype::TimeLoop*>(loop);
  if (int rc = reducer_failed(exahype::time_loop_reducer(l))) return rc;
  const bool pending = exahype::peer_reducer_pending(exahype::time_loop_reducer(l));
  cudaError_t err = exahype::time_loop_flush(l, static_cast<cudaStream_t>(stream));
  if (err != cudaSuccess) return cuda_fail(err, "loop_consume_kernel launch (flush)");
  if (pending) g_launches.fetch_add(1);
  return EXAHYPE_OK;
}

int exahype_cuda_time_loop_history(void* loop, int64_t first_step, int64_t count, void* out) {
  if (!loop || !out) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "null loop / out");
  exahype::TimeLoop* l = static_cast<exahype::TimeLoop*>(loop);
  cudaError_t err = exahype::time_loop_history(l, first_step, count, out);
  if (err == cudaErrorInvalidValue)
    return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "history range [%lld, %lld) outside the recorded steps / capacity",
                (long long)first_step, (long long)(first_step + count));
  if (err != cudaSuccess) return cuda_fail(err, "time loop history");
  return reducer_failed(exahype::time_loop_reducer(l));
}

int64_t exahype_cuda_time_loop_steps(void* loop) {
  return loop ? exahype::time_loop_steps(static_cast<exahype::TimeLoop*>(loop)) : 0;
}

int exahype_cuda_time_loop_dt_device(void* loop, void** dt_device) {
  if (!loop || !dt_device) return fail(EXAHYPE_ERR_INVALID_ARGUMENT, "null loop / dt_device");
  *dt_device = exahype::time_loop_dt_device(static_cast<exahype::TimeLoop*>(loop));
  return EXAHYPE_OK;
}

int exahype_cuda_time_loop_destroy(void* loop) {
  exahype::time_loop_destroy(static_cast<exahype::TimeLoop*>(loop));
  return EXAHYPE_OK;
}

#pragma GCC visibility pop
}  // extern "C"
