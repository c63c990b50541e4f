// Benchmark / test input on the device: the counter-based admissible synthetic state of SURVEY.md section 8d.
// Bit-identical to the test oracle's generator (oracle/fv_rusanov_oracle.c: fvo_synth_cell) without calling it: every
// slot of the flat AoS batch is a SplitMix64 hash of its own index, so any shard can be generated anywhere without
// communication and a 1.3 GB shard never crosses PCIe.  Compiled with -fmad=false like the kernels (no contraction).
#include "../../include/exahype_cuda.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ double u01(uint64_t idx, uint64_t seed) {
  uint64_t z = (idx + seed) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

template <typename T>
__global__ void fill_synthetic_kernel(T* __restrict__ q, long long first_cell, long long n_cells, unsigned long long seed,
                                      int model, int dim, int nv) {
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n_cells; c += (long long)gridDim.x * blockDim.x) {
    const uint64_t base = (uint64_t)(first_cell + c) * (uint64_t)nv;
    T* out = q + c * nv;
    if (model == EXAHYPE_MODEL_EULER) {
      const double rho = 1.0 + u01(base, seed);
      double ke = 0.0;
      for (int k = 0; k < dim; ++k) {
        const double vel = u01(base + 1 + k, seed) - 0.5;
        out[1 + k] = (T)(rho * vel);
        ke = ke + vel * vel;
      }
      const double p = 1.0 + u01(base + dim + 1, seed);
      out[0] = (T)rho;
      out[dim + 1] = (T)(p / (1.4 - 1.0) + 0.5 * rho * ke);
      for (int v = dim + 2; v < nv; ++v) out[v] = (T)u01(base + v, seed);
    } else {
      const double hgt = 1.0 + u01(base, seed);
      out[0] = (T)hgt;
      out[1] = (T)(hgt * (0.2 * (u01(base + 1, seed) - 0.5)));
      out[2] = (T)(hgt * (0.2 * (u01(base + 2, seed) - 0.5)));
      for (int v = 3; v < nv; ++v) out[v] = (T)(0.1 * u01(base + v, seed));
    }
  }
}

}  // namespace

namespace exahype {
cudaError_t fill_synthetic(const exahype_fv_config* cfg, void* q, long long first_cell, long long n_cells,
                           unsigned long long seed, cudaStream_t stream) {
  if (n_cells <= 0) return cudaSuccess;
  const int nv = cfg->n_real + cfg->n_aux;
  const int block = 256;
  long long blocks = (n_cells + block - 1) / block;
  const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
  if (cfg->dtype == EXAHYPE_DTYPE_F64)
    fill_synthetic_kernel<double><<<grid, block, 0, stream>>>(static_cast<double*>(q), first_cell, n_cells, seed, cfg->model, cfg->dim, nv);
  else
    fill_synthetic_kernel<float><<<grid, block, 0, stream>>>(static_cast<float*>(q), first_cell, n_cells, seed, cfg->model, cfg->dim, nv);
  return cudaGetLastError();
}
}  // namespace exahype
