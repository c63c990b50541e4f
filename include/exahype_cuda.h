/*
 * libexahype_cuda.so -- C ABI of the B200-native batched stateless finite-volume Rusanov patch update.
 *
 * This is the drop-in boundary for ONE path of xdslproject/ExaHyPE: the generated kernel
 *     void time_step(double* Q, double dt);                 (reference: Unit test/test.h:3,
 *                                                            body Unit test/test.cpp:3-111,
 *                                                            emitted by exahype/printers/CPPPrinter.py:48-102)
 * together with the user physics it links against
 *     void   Flux(const double* Q, int normal, double* F);  (reference: Unit test/Functions.h:2)
 *     double maxEigenvalue(const double* Q, int normal);    (reference: Unit test/Functions.h:3)
 *     double max(double* a, double* b);                     (reference: Unit test/Functions.h:4)
 *
 * Plain pointers and sizes only; no torch / C++ types.  Device entry points take DEVICE pointers and a
 * cudaStream_t passed as void*; they are asynchronous and allocate nothing.  Every function returns
 * EXAHYPE_OK (0) or a negative error code; exahype_cuda_last_error() gives the thread-local message.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with EXAHYPE_ERR_CUDA.
 *
 * Data layout (reference: exahype/printers/CPPPrinter.py:247-261, Unit test/test.cpp:15):
 *     Q[patch][i][j]([k])[var]   AoS, haloed, side S = patch_size + 2*halo, var < n_real + n_aux,
 *     `i` slowest spatial index and paired with normal = 0.
 */
#ifndef EXAHYPE_CUDA_H
#define EXAHYPE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EXAHYPE_CUDA_ABI_VERSION 2

enum {
  EXAHYPE_OK = 0,
  EXAHYPE_ERR_INVALID_ARGUMENT = -1, /* mirrors KernelBuilder.viable() (reference KernelBuilder.py:41-48) + pointer/size checks */
  EXAHYPE_ERR_NO_INSTANTIATION = -2, /* no committed kernel for this (model, dim, patch, halo, vars, dtype): generate one with CUDAPrinter */
  EXAHYPE_ERR_CUDA = -3,             /* CUDA runtime error (message carries cudaGetErrorString) */
  EXAHYPE_ERR_NCCL = -4,
  EXAHYPE_ERR_UNAVAILABLE = -5,      /* optional component missing (e.g. libnccl not loadable) */
  EXAHYPE_ERR_TIMEOUT = -6           /* a peer never arrived at an exchange: results derived from it are NaN (sticky) */
};

/* physics families with committed hand-written functors (csrc/physics.cuh) */
/* EXAHYPE_MODEL_SWE_SOURCE: shallow water with the bathymetry source term: cell variables
 * (h, hu, hv | b, db/dx, db/dy), n_real = 3, n_aux = 3; one more statement after the dissipation,
 * Q_copy = Q_copy + dt*S with S = (0, -g h db/dx, -g h db/dy) of the original state.  The reference has only the solver
 * signature sourceTerm(Q, x, h, t, dt, S) ("Unit test/correctness_test.cpp":16-23); the definition is this repository's. */
enum { EXAHYPE_MODEL_EULER = 0, EXAHYPE_MODEL_SWE = 1, EXAHYPE_MODEL_SWE_SOURCE = 2 };
enum { EXAHYPE_DTYPE_F64 = 0, EXAHYPE_DTYPE_F32 = 1 };

/* flags */
enum {
  /* Rusanov dissipation on all n_real unknowns (what `struct=True` at examples/Batched_stateless.py:33 intends).
   * Default (flag clear) is the reference-emitted behaviour: variable 0 only (Unit test/test.cpp:81,90). */
  EXAHYPE_FLAG_DISSIPATION_ALL = 1u << 0,
  /* q_out is un-haloed: q_out[patch][P]^dim[n_real+n_aux] (ExaHyPE2's QOut convention, reference
   * exahype/printers/CPPPrinter.py:255-258, examples/kernel-generator.py:11-12).  Default: q_out has the haloed
   * layout of q_in and only interior cells are written (q_out == q_in gives the reference's in-place update). */
  EXAHYPE_FLAG_OUTPUT_UNHALOED = 1u << 1,
  /* do not reset *lambda_max to 0 before the launch: fold this launch into the running maximum */
  EXAHYPE_FLAG_LAMBDA_ACCUMULATE = 1u << 2,
  /* 3-D shapes have two kernels with identical results: plane marching (default) and thread-per-cell (this flag).
   * Ignored where only one kernel exists. */
  EXAHYPE_FLAG_KERNEL_CELL = 1u << 3,
  /* Permission to leave the reference's bit pattern: fused multiply-adds and a branch-free reciprocal / square root
   * (valid for normal-range densities / heights).  Results stay within 1e-12 relative of the reference arithmetic -- the
   * bound BASELINE.json states -- but are no longer bitwise equal to it.  Committed for the headline shapes (8^3 Euler
   * fp64, 16^2 Euler fp64, 32^2 shallow water fp64 / fp32); other shapes keep the reference arithmetic.  The kernels
   * execute ~14 % fewer instructions, which is what bounds them once the board runs power-limited. */
  EXAHYPE_FLAG_FAST_ARITHMETIC = 1u << 4,
  /* With EXAHYPE_FLAG_OUTPUT_UNHALOED: q_out carries the unknowns only, q_out[patch][P]^dim[n_real].  The step never
   * changes the auxiliary variables (every update statement of the reference runs over the unknowns, "Unit test/
   * test.cpp":60-95; the copy-back at :96-103 passes the auxiliary values through as they came in), and an
   * un-haloed QOut that repeats them costs their bytes a second time: 13 % of the DRAM traffic of a 32x32 shallow-water
   * batch, which runs at the HBM wall.  No effect for n_aux == 0.  Committed for the row-marching kernel of the
   * shallow-water families (dense batches: exahype_cuda_fv_step / _allreduce / _time_loop); EXAHYPE_ERR_NO_INSTANTIATION
   * elsewhere. */
  EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY = 1u << 5
};

typedef struct {
  int32_t model;      /* EXAHYPE_MODEL_* */
  int32_t dtype;      /* EXAHYPE_DTYPE_* */
  int32_t dim;        /* 2 | 3 */
  int32_t patch_size; /* P */
  int32_t halo;       /* h >= 1 */
  int32_t n_real;
  int32_t n_aux;
  uint32_t flags;     /* EXAHYPE_FLAG_* */
} exahype_fv_config;

int exahype_cuda_version(void);
const char* exahype_cuda_last_error(void);
/* number of visible CUDA devices (0 on a host without a GPU; never fails) */
int exahype_cuda_device_count(void);
/* 1 if a committed instantiation exists for cfg (ignores flags), else 0 */
int exahype_cuda_fv_supported(const exahype_fv_config* cfg);
/* Fills up to `capacity` configs with the committed instantiations; returns how many exist. */
int exahype_cuda_fv_list(exahype_fv_config* out, int capacity);

/*
 * One patch-update step over a batch resident in device memory: replaces `time_step(Q, dt)`
 * (reference Unit test/test.h:3) for n_patches patches at once.
 *   q_in          device, haloed AoS batch, n_patches * S^dim * (n_real+n_aux) elements of cfg->dtype, 16-byte aligned
 *   q_out         device; == q_in for the reference's in-place semantics, or a second buffer (haloed or un-haloed per flags)
 *   dt            time step (converted to cfg->dtype)
 *   lambda_patch  device, nullable: n_patches values, max over interior cells and directions of maxEigenvalue(Q_in)
 *   lambda_max    device, nullable: 1 value, max over the batch (reset to 0 first unless EXAHYPE_FLAG_LAMBDA_ACCUMULATE)
 *   stream        cudaStream_t
 */
int exahype_cuda_fv_step(const exahype_fv_config* cfg, const void* q_in, void* q_out, int64_t n_patches,
                         double dt, void* lambda_patch, void* lambda_max, void* stream);

/*
 * The ExaHyPE2 `CellData` form of the same step (reference examples/kernel-generator.py:8-19: `patchData` of type
 * `::exahype2::CellData&` with members QIn, QOut, dt; the production boundary of Peano's
 * timeStepWithRusanovBatchedStateless, "Unit test/correctness_test.cpp":138-163): the batch is not one dense array but a
 * list of patches, each with its own pointers and its own time step.
 *   q_in            device array of n_patches device pointers, each to one haloed patch S^dim * (n_real+n_aux), 16-byte aligned
 *   q_out           device array of n_patches device pointers; q_out[p] == q_in[p] updates patch p in place (haloed),
 *                   with EXAHYPE_FLAG_OUTPUT_UNHALOED each points to P^dim * (n_real+n_aux) values (CellData::QOut)
 *   dt              device array of n_patches time steps of cfg->dtype (CellData::dt); null: the scalar `dt` argument
 *   max_eigenvalue  device array of n_patches, nullable (CellData::maxEigenvalue): as lambda_patch of exahype_cuda_fv_step
 *   cell_centre     device array of n_patches * dim values (CellData::cellCentre), nullable: 0
 *   cell_size       device array of n_patches * dim values (CellData::cellSize), nullable: 1
 *   t               device array of n_patches time stamps (CellData::t), nullable: 0
 * cell_centre / cell_size / t reach functors declared with ExaHyPE2's solver signature flux(Q, x, h, t, dt, normal, F)
 * (kernel-generator.py:37-39; units generated by CUDAPrinter from such a declaration); the committed Euler and
 * shallow-water families do not depend on position or time and ignore them.
 */
typedef struct {
  int64_t n_patches;
  const void* const* q_in;
  void* const* q_out;
  const void* dt;
  void* max_eigenvalue;
  const void* cell_centre;
  const void* cell_size;
  const void* t;
} exahype_cell_data;

int exahype_cuda_fv_step_cell_data(const exahype_fv_config* cfg, const exahype_cell_data* cells, double dt,
                                   void* lambda_max, void* stream);

/* Named entry points for the committed headline instantiations (same semantics as exahype_cuda_fv_step). */
int exahype_cuda_fv_step_euler_2d_f64(const double* q_in, double* q_out, int64_t n_patches, int patch_size, int halo,
                                      int n_aux, double dt, double* lambda_patch, double* lambda_max,
                                      unsigned flags, void* stream);
int exahype_cuda_fv_step_euler_3d_f64(const double* q_in, double* q_out, int64_t n_patches, int patch_size, int halo,
                                      int n_aux, double dt, double* lambda_patch, double* lambda_max,
                                      unsigned flags, void* stream);
int exahype_cuda_fv_step_swe_2d_f64(const double* q_in, double* q_out, int64_t n_patches, int patch_size, int halo,
                                    int n_aux, double dt, double* lambda_patch, double* lambda_max,
                                    unsigned flags, void* stream);
int exahype_cuda_fv_step_swe_2d_f32(const float* q_in, float* q_out, int64_t n_patches, int patch_size, int halo,
                                    int n_aux, float dt, float* lambda_patch, float* lambda_max,
                                    unsigned flags, void* stream);

/*
 * The reference's own call shape, `time_step(Q, dt)` on HOST memory (reference Unit test/correctness_test.cpp:195):
 * copies the batch to the device in chunks, updates it there and copies the interior back, with copies and kernels
 * overlapped on internal streams.  q_host is updated in place (haloed) or q_out_host receives the result
 * (layout per flags; may equal q_host when haloed).  lambda_max_host (nullable) receives the batch maximum.
 * Pinned host memory gives full PCIe bandwidth; pageable memory works.  Synchronous: returns when done.
 * Thread-safe: the staging ring is per device (the caller's current device); host threads driving different GPUs run
 * concurrently, two calls for the same device serialise.
 */
int exahype_cuda_time_step_host(const exahype_fv_config* cfg, const void* q_host, void* q_out_host,
                                int64_t n_patches, double dt, void* lambda_patch_host, void* lambda_max_host);
/* Releases the staging buffers and streams cached by exahype_cuda_time_step_host for the current device. */
int exahype_cuda_host_pipeline_release(void);
/* Chunk size (patches) and number of in-flight chunks used by exahype_cuda_time_step_host; 0 keeps the default. */
int exahype_cuda_host_pipeline_configure(int64_t chunk_patches, int depth);

/*
 * Synthetic admissible input of the benchmark generated on the device: n_cells haloed cells
 * starting at global cell index first_cell (= first_patch * S^dim), every slot a SplitMix64 hash of its own flat index, so
 * shards generated on different GPUs are the slices of one global batch.  New; the reference has no benchmark input.
 */
#define EXAHYPE_SYNTHETIC_SEED 20240601ull
int exahype_cuda_fill_synthetic(const exahype_fv_config* cfg, void* q, int64_t first_cell, int64_t n_cells,
                                uint64_t seed, void* stream);

/* Number of kernels launched by this library in this process since load (for bench.py's gpu_launches). */
int64_t exahype_cuda_launch_count(void);
/* Kernel geometry picked for cfg on the current device: grid, block, dynamic smem bytes, patches per tile. */
int exahype_cuda_fv_launch_info(const exahype_fv_config* cfg, int64_t n_patches, int* grid, int* block,
                                int* smem_bytes, int* patches_per_tile);

/*
 * Global admissible time step across GPUs: one NCCL all-reduce(max) of a single scalar
 * (new; the reference has no distributed code).  One communicator per process/GPU.
 *   exahype_cuda_nccl_unique_id  writes a 128-byte ncclUniqueId (rank 0 creates it, the host plumbing broadcasts it)
 *   exahype_cuda_comm_init       ncclCommInitRank on the current device; *comm receives an opaque handle
 *   exahype_cuda_allreduce_max   in-place max over ranks of count values of dtype at device pointer `values`
 */
int exahype_cuda_nccl_unique_id(void* id128);
int exahype_cuda_comm_init(void** comm, const void* id128, int world_size, int rank);
int exahype_cuda_comm_destroy(void* comm);
int exahype_cuda_allreduce_max(void* comm, void* values, int64_t count, int dtype, void* stream);

/*
 * The same reduction without NCCL: a one-shot all-reduce(max) of ONE scalar over NVLink peer memory (each rank's
 * mailbox is mapped into every peer through CUDA IPC; one tiny stream-ordered kernel per step stores the local value into
 * every peer's mailbox and spins on its own).  Latency ~ one NVLink round trip instead of an NCCL launch.  Collective:
 * every rank calls allreduce_max the same number of times.  One process per GPU, all on one node.
 *   create        allocates this rank's mailbox on the current device
 *   local_handle  writes the 64-byte cudaIpcMemHandle_t of the mailbox (the host plumbing all-gathers them)
 *   connect       maps the world_size * 64 bytes of gathered handles (own entry ignored)
 *   allreduce_max in place on the device scalar `value` of dtype, asynchronous on stream
 *   status        *flag = 1 if a wait timed out (a peer never arrived).  The flag lives in host-mapped memory: reading it
 *                 synchronises nothing, and every later call on the reducer (or a time loop built on it) fails with
 *                 EXAHYPE_ERR_TIMEOUT.  The waiting side's result is NaN, never a silently rank-local value.
 *   set_timeout   how long a wait spins before it gives up (default 10 s)
 *   connect_local test / several-ranks-per-GPU form of connect: all `world_size` reducers live in this process on one device
 *   enable_trace  keep globaltimer stamps of the last `capacity` exchanges (see csrc/peer_mail.cuh FV_TRACE_*);
 *   read_trace    host copy of count rows of 8 uint64 starting at exchange first_seq (synchronises the device)
 */
int exahype_cuda_peer_reducer_create(void** reducer, int world_size, int rank);
int exahype_cuda_peer_reducer_local_handle(void* reducer, void* handle64);
int exahype_cuda_peer_reducer_connect(void* reducer, const void* all_handles);
int exahype_cuda_peer_reducer_allreduce_max(void* reducer, void* value, int dtype, void* stream);
int exahype_cuda_peer_reducer_status(void* reducer, int* flag);
int exahype_cuda_peer_reducer_set_timeout(void* reducer, double seconds);
int exahype_cuda_peer_reducer_connect_local(void* const* reducers, int world_size);
int exahype_cuda_peer_reducer_enable_trace(void* reducer, int capacity);
int exahype_cuda_peer_reducer_read_trace(void* reducer, uint64_t first_seq, int count, uint64_t* out);
/*
 * exahype_cuda_fv_step followed by exahype_cuda_peer_reducer_allreduce_max(lambda_max), as ONE launch where the shape's
 * kernel supports it (the warp-per-patch kernel of 8x8x8 patches): the last warp of the grid to finish exchanges the
 * device's maximum with every peer from the kernel's own epilogue, so no second kernel sits between two steps.  Other
 * shapes (and empty shards) fall back to the two launches.  Same arguments and flags as exahype_cuda_fv_step;
 * lambda_max is required and holds the maximum over ALL ranks afterwards.  Collective: counts as one allreduce_max call.
 * The reference has no counterpart (its kernel is serial).
 */
int exahype_cuda_fv_step_allreduce(const exahype_fv_config* cfg, void* reducer, const void* q_in, void* q_out,
                                   int64_t n_patches, double dt, void* lambda_patch, void* lambda_max, void* stream);
int exahype_cuda_peer_reducer_destroy(void* reducer);

/*
 * Device-resident time loop: the global admissible time step of SURVEY.md section 8e, produced AND consumed on the
 * device.  Step k+1 advances with dt = cfl_dx / max over all ranks of lambda_max(step k), where lambda_max(step k) is the
 * largest eigenvalue of step k's input state (a8).  Nothing crosses to the host between steps, there is no memset and
 * no second kernel between two steps of the warp-per-patch kernel, and the exchange is split-phase: the last warp of
 * step k's grid stores this device's maximum into every peer's mailbox (NVLink peer memory, csrc/peer_mail.cuh) and
 * the launch ends; the warps of step k+1 read the mailbox right before their first use of dt, so waiting for the
 * slowest rank overlaps step k+1's start-up.  max is exact and all ranks divide the same two numbers: an N-GPU run
 * takes bit for bit the time steps of the 1-GPU run.  The reference has no counterpart: its `time_step(Q, dt)` takes
 * dt by value ("Unit test/test.h":3) and Peano derives the next dt on the host.
 *   create    dtype EXAHYPE_DTYPE_*; reducer: a connected peer reducer (all ranks step in lockstep), or NULL for this
 *             device alone; cfl_dx = CFL number x cell size; dt0 = time step of the first step;
 *             history_capacity = steps remembered (0: 4096)
 *   step      as exahype_cuda_fv_step, with dt and lambda_max owned by the loop.  Shapes whose kernel has no exchange
 *             epilogue (and empty shards) run the same protocol with a one-warp kernel behind the patch kernel.
 *             Collective across the reducer's ranks: every rank calls it once per step.
 *   flush     consumes the exchange still in flight (a one-warp kernel): afterwards dt_device holds the NEXT step's dt
 *   history   host copy of entries [first_step, first_step + count), 4 values of dtype each: {dt used by the step, the
 *             global maximum it was derived from (0: none, dt0 was used), this device's lambda_max of the step, 0}.
 *             Entry `steps` (after flush) holds {next dt, last global maximum}.  Synchronises the device.
 *   steps     steps launched so far;  dt_device: device scalar with the most recently established dt
 */
int exahype_cuda_time_loop_create(void** loop, int dtype, void* reducer, double cfl_dx, double dt0,
                                  int64_t history_capacity);
int exahype_cuda_fv_step_time_loop(const exahype_fv_config* cfg, void* loop, const void* q_in, void* q_out,
                                   int64_t n_patches, void* lambda_patch, void* stream);
int exahype_cuda_time_loop_flush(void* loop, void* stream);
int exahype_cuda_time_loop_history(void* loop, int64_t first_step, int64_t count, void* out);
int64_t exahype_cuda_time_loop_steps(void* loop);
int exahype_cuda_time_loop_dt_device(void* loop, void** dt_device);
int exahype_cuda_time_loop_destroy(void* loop);

#ifdef __cplusplus
}
#endif
#endif /* EXAHYPE_CUDA_H */
