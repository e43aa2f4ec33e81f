// FP64 tensor-core (mma.sync.m8n8k4.f64) peak of one B200 against the FP64 FMA pipe.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a dmma_peak.cu -o dmma_peak
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void __launch_bounds__(256) dmma_chains(double* out, int iters, double a, double b) {
  double c[NACC][2];
  for (int k = 0; k < NACC; ++k) c[k][0] = c[k][1] = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) dmma(c[k], a, b);
  }
  double s = 0;
  for (int k = 0; k < NACC; ++k) s += c[k][0] + c[k][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dfma_chains(double* out, int iters, double a, double b) {
  double x[16];
  for (int k = 0; k < 16; ++k) x[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0;
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* sink; cudaMalloc(&sink, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int ctas = 148 * 8, iters = 4096;
  auto run = [&](const char* name, auto launch, double flops) {
    launch(); launch();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    printf("%-24s %8.3f ms  %7.2f TFLOP/s\n", name, best, flops / (best * 1e-3) / 1e12);
  };
  run("DFMA 16 chains", [&] { dfma_chains<<<ctas, 256>>>(sink, iters, 0.999999, 1e-7); }, 2.0 * 16 * iters * 256.0 * ctas);
  // one warp-level m8n8k4 = 8*8*4 FMAs = 512 flops; per thread per mma: 512/32 = 16 flops
  run("DMMA m8n8k4 4 acc", [&] { dmma_chains<4><<<ctas, 256>>>(sink, iters, 0.999999, 1e-7); }, 16.0 * 4 * iters * 256.0 * ctas);
  run("DMMA m8n8k4 8 acc", [&] { dmma_chains<8><<<ctas, 256>>>(sink, iters, 0.999999, 1e-7); }, 16.0 * 8 * iters * 256.0 * ctas);
  run("DMMA m8n8k4 16 acc", [&] { dmma_chains<16><<<ctas, 256>>>(sink, iters, 0.999999, 1e-7); }, 16.0 * 16 * iters * 256.0 * ctas);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
