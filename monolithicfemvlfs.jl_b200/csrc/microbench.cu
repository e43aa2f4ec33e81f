// microbench.cu - measured roofline denominators of THIS device for the kernels of the path that
// are not HBM-bound: the FP64 FMA rate (general-quadrature kernel integrate_cells, the Schur
// updates of the multifrontal LU).  SURVEY.md §8(d): "report against a measured FP64 FMA peak".
#include "common.cuh"

namespace {
// 16 independent DFMA chains per thread, no memory traffic inside the loop
__global__ void __launch_bounds__(256) dfma_chains(double* __restrict__ out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // never true: keeps the chains alive
}
}  // namespace

// TFLOP/s of dependent-free FP64 FMAs on every SM (2 flops per FMA), best of `reps` launches
double measure_fp64_fma_tflops(int nsm, int reps, cudaStream_t s) {
  DevBuf<double> sink;
  const int ctas = nsm * 8, iters = 4096;
  sink.alloc((size_t)ctas * 256);
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  dfma_chains<<<ctas, 256, 0, s>>>(sink.p, iters, 0.999999, 1e-7);      // warm-up (clocks)
  dfma_chains<<<ctas, 256, 0, s>>>(sink.p, iters, 0.999999, 1e-7);
  double best = 0.0;
  for (int r = 0; r < reps; ++r) {
    CUDA_TRY(cudaEventRecord(e0, s));
    dfma_chains<<<ctas, 256, 0, s>>>(sink.p, iters, 0.999999, 1e-7);
    CUDA_TRY(cudaEventRecord(e1, s));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16.0 * iters * 256.0 * ctas;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  CUDA_TRY(cudaGetLastError());
  return best;
}
