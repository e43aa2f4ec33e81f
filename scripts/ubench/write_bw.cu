// write-bandwidth microbenchmark: how fast can one B200 stream 16-byte stores (no reads)?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a write_bw.cu -o write_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
template <int MODE>
__global__ void fill(double2* __restrict__ out, size_t n, double2 v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (MODE == 0) out[i] = v;
    else if (MODE == 1) __stcs(out + i, v);
    else if (MODE == 2) __stcg(out + i, v);
    else __stwt(out + i, v);
  }
}
__global__ void copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
// rows of `len` items copied from a small (L2 resident) template of `tn` items
__global__ void expand(const double2* __restrict__ tm, size_t tn, double2* __restrict__ out, size_t n, int len) {
  const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (size_t r = w; r * len < n; r += nw) {
    const size_t s = (r * 2654435761ull) % (tn - len);
    for (int k = lane; k < len; k += 32) __stcs(out + r * len + k, __ldg(tm + s + k));
  }
}
// mode 1: `grp` consecutive rows share a template row (L1 hits), rows written in order
// mode 2: same sharing, destination rows permuted pseudo-randomly (bijective multiplicative hash on the row index)
__global__ void expand2(const double2* __restrict__ tm, size_t tn, double2* __restrict__ out, size_t nrows, int len, int grp, int perm) {
  const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  // a warp takes `grp` consecutive rows: same template
  for (size_t g = w; g * grp < nrows; g += nw) {
    const size_t s = (g * 2654435761ull) % (tn - len);
    for (int q = 0; q < grp; ++q) {
      size_t r = g * grp + q;
      if (r >= nrows) break;
      if (perm) r = (r * 7919ull + 12345ull) % nrows;      // 7919 prime: bijection when gcd(7919, nrows) = 1
      for (int k = lane; k < len; k += 32) __stcs(out + r * len + k, __ldg(tm + s + k));
    }
  }
}
int main() {
  const size_t n = 77528672;      // Y1 nnz
  double2 *a, *b;
  cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16);
  cudaMemset(a, 1, n * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto f, double bytes) {
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
    printf("%-28s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms * 1e-6);
  };
  const double2 v = make_double2(1.0, 2.0);
  for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32, 148 * 64}) {
    char nm[64];
    snprintf(nm, 64, "fill st       grid %d", g); timeit(nm, [&] { fill<0><<<g, 256>>>(b, n, v); }, n * 16.0);
    snprintf(nm, 64, "fill st.cs    grid %d", g); timeit(nm, [&] { fill<1><<<g, 256>>>(b, n, v); }, n * 16.0);
    snprintf(nm, 64, "fill st.cg    grid %d", g); timeit(nm, [&] { fill<2><<<g, 256>>>(b, n, v); }, n * 16.0);
    snprintf(nm, 64, "fill st.wt    grid %d", g); timeit(nm, [&] { fill<3><<<g, 256>>>(b, n, v); }, n * 16.0);
  }
  timeit("cudaMemsetAsync", [&] { cudaMemsetAsync(b, 0, n * 16); }, n * 16.0);
  timeit("copy (read+write bytes)", [&] { copy<<<148 * 16, 256>>>(a, b, n); }, n * 32.0);
  timeit("cudaMemcpyAsync D2D (r+w)", [&] { cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice); }, n * 32.0);
  for (int len : {32, 64, 96, 128, 256}) {
    char nm[64];
    snprintf(nm, 64, "expand len %d (tmpl 32 MB)", len);
    timeit(nm, [&] { expand<<<148 * 16, 256>>>(a, 2000000, b, n, len); }, n * 16.0);
  }
  for (int len : {64, 96, 128}) for (int perm : {0, 1}) for (int grp : {1, 8, 32}) {
    char nm[64];
    snprintf(nm, 64, "expand2 len %d grp %d perm %d", len, grp, perm);
    const size_t nrows = n / len;
    timeit(nm, [&] { expand2<<<148 * 16, 256>>>(a, 2000000, b, nrows, len, grp, perm); }, nrows * (double)len * 16.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
