// Microbenchmark: do warp shuffles share the shared-memory data pipe with LDS/STS on sm_100a?
// Times (cycles per warp-instruction per SM) of: LDS.64 only, SHFL.32 only, both interleaved, at 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, int iters) {
  __shared__ double s[32 * 40];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 40; i += blockDim.x) s[i] = i;
  __syncthreads();
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int b0 = lane, b1 = lane + 1, b2 = lane + 2, b3 = lane + 3;
  const double* p = s + warp * 32 + lane;
  for (int it = 0; it < iters; ++it) {
    if (MODE & 1) {
      double t0, t1, t2, t3;
      const unsigned addr = (unsigned)__cvta_generic_to_shared(p) + ((it & 3) << 3);
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t0) : "r"(addr) : "memory");
      asm volatile("ld.shared.f64 %0, [%1+256];" : "=d"(t1) : "r"(addr) : "memory");
      asm volatile("ld.shared.f64 %0, [%1+512];" : "=d"(t2) : "r"(addr) : "memory");
      asm volatile("ld.shared.f64 %0, [%1+768];" : "=d"(t3) : "r"(addr) : "memory");
      a0 += t0; a1 += t1; a2 += t2; a3 += t3;
    }
    if (MODE & 2) {
      b0 = __shfl_sync(0xffffffffu, b0, (lane + 1) & 31);
      b1 = __shfl_sync(0xffffffffu, b1, (lane + 1) & 31);
      b2 = __shfl_sync(0xffffffffu, b2, (lane + 1) & 31);
      b3 = __shfl_sync(0xffffffffu, b3, (lane + 1) & 31);
      b0 = __shfl_sync(0xffffffffu, b0, (lane + 31) & 31);
      b1 = __shfl_sync(0xffffffffu, b1, (lane + 31) & 31);
      b2 = __shfl_sync(0xffffffffu, b2, (lane + 31) & 31);
      b3 = __shfl_sync(0xffffffffu, b3, (lane + 31) & 31);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + b0 + b1 + b2 + b3;
}

template <int MODE>
float run(double* out, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, 512>>>(out, iters);
  cudaEventRecord(a);
  k<MODE><<<148, 512>>>(out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  double* out; cudaMalloc(&out, 148 * 512 * 8);
  const int iters = 20000;
  float lds = run<1>(out, iters), shf = run<2>(out, iters), both = run<3>(out, iters);
  // per SM: 16 warps x iters x (4 LDS.64 | 8 SHFL)
  const double clk = 1.965e9;
  printf("LDS.64 only : %.3f ms  -> %.2f cycles per warp-LDS.64 per SM\n", lds, lds * 1e-3 * clk / (16.0 * iters * 4));
  printf("SHFL only   : %.3f ms  -> %.2f cycles per warp-SHFL per SM\n", shf, shf * 1e-3 * clk / (16.0 * iters * 8));
  printf("both        : %.3f ms  (sum %.3f, max %.3f)\n", both, lds + shf, lds > shf ? lds : shf);
  return 0;
}
