// spmv.cu - CSR sparse matrix-vector products, real and complex FP64 (sm_100a).
//
// Replaces the products Gridap/SparseArrays perform inside the Newmark step
// (src/Khabakhpasheva_time_domain.jl:258-266, src/Periodic_Beam.jl:107-116) and
// inside every Krylov iteration of the solve that stands in for `LUSolver()`
// (src/Yago_freq_domain.jl:171).  HBM-bound: nnz*(sT+4) + 4(N+1) + 2*N*sT bytes.
// A group of TPR lanes owns a row; values are streamed with 128-bit loads, x is
// gathered through the read-only path (it lives in the 126 MB L2), the row sum is a
// shuffle reduction.
#include "common.cuh"

namespace {

struct cplx { double x, y; };

__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ double2 ld_stream(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }

template <int TPR>
__global__ void __launch_bounds__(256)
csr_spmv_real(const int* __restrict__ rowptr, const int* __restrict__ colind,
              const double* __restrict__ vals, const double* __restrict__ x,
              double* __restrict__ y, long long nrows, double alpha, double beta, const int* __restrict__ rows) {
  const int lane = threadIdx.x & (TPR - 1);
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / TPR;
  const long long nrg = ((long long)gridDim.x * blockDim.x) / TPR;
  for (long long q = gid; q < nrows; q += nrg) {
    const long long r = rows ? rows[q] : q;          // optional row list (interior / boundary rows of a rank)
    const int p0 = rowptr[r], p1 = rowptr[r + 1];
    double s0 = 0, s1 = 0;
    int p = p0 + lane;
    for (; p + TPR < p1; p += 2 * TPR) {
      int c0 = ld_stream(colind + p), c1 = ld_stream(colind + p + TPR);
      double v0 = ld_stream(vals + p), v1 = ld_stream(vals + p + TPR);
      s0 = fma(v0, __ldg(x + c0), s0);
      s1 = fma(v1, __ldg(x + c1), s1);
    }
    if (p < p1) s0 = fma(ld_stream(vals + p), __ldg(x + ld_stream(colind + p)), s0);
    double s = s0 + s1;
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, TPR);
    if (lane == 0) y[r] = beta == 0.0 ? alpha * s : fma(alpha, s, beta * y[r]);
  }
}

template <int TPR>
__global__ void __launch_bounds__(256)
csr_spmv_cplx(const int* __restrict__ rowptr, const int* __restrict__ colind,
              const double2* __restrict__ vals, const double2* __restrict__ x,
              double2* __restrict__ y, long long nrows, double2 alpha, double2 beta, const int* __restrict__ rows) {
  const int lane = threadIdx.x & (TPR - 1);
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / TPR;
  const long long nrg = ((long long)gridDim.x * blockDim.x) / TPR;
  for (long long q = gid; q < nrows; q += nrg) {
    const long long r = rows ? rows[q] : q;          // optional row list (interior / boundary rows of a rank)
    const int p0 = rowptr[r], p1 = rowptr[r + 1];
    double sr0 = 0, si0 = 0, sr1 = 0, si1 = 0;
    int p = p0 + lane;
    for (; p + TPR < p1; p += 2 * TPR) {
      int c0 = ld_stream(colind + p), c1 = ld_stream(colind + p + TPR);
      double2 v0 = ld_stream(vals + p), v1 = ld_stream(vals + p + TPR);
      double2 x0 = __ldg(x + c0), x1 = __ldg(x + c1);
      sr0 = fma(v0.x, x0.x, sr0); sr0 = fma(-v0.y, x0.y, sr0);
      si0 = fma(v0.x, x0.y, si0); si0 = fma(v0.y, x0.x, si0);
      sr1 = fma(v1.x, x1.x, sr1); sr1 = fma(-v1.y, x1.y, sr1);
      si1 = fma(v1.x, x1.y, si1); si1 = fma(v1.y, x1.x, si1);
    }
    if (p < p1) {
      double2 v0 = ld_stream(vals + p);
      double2 x0 = __ldg(x + ld_stream(colind + p));
      sr0 = fma(v0.x, x0.x, sr0); sr0 = fma(-v0.y, x0.y, sr0);
      si0 = fma(v0.x, x0.y, si0); si0 = fma(v0.y, x0.x, si0);
    }
    double sr = sr0 + sr1, si = si0 + si1;
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) {
      sr += __shfl_down_sync(0xffffffffu, sr, o, TPR);
      si += __shfl_down_sync(0xffffffffu, si, o, TPR);
    }
    if (lane == 0) {
      double2 a;
      a.x = alpha.x * sr - alpha.y * si;
      a.y = alpha.x * si + alpha.y * sr;
      if (beta.x != 0.0 || beta.y != 0.0) {
        double2 yo = y[r];
        a.x += beta.x * yo.x - beta.y * yo.y;
        a.y += beta.x * yo.y + beta.y * yo.x;
      }
      y[r] = a;
    }
  }
}

}  // namespace

// y = alpha*A*x + beta*y ; alpha/beta given as (re,im) pairs (im ignored when real)
void launch_spmv_rows(const int* rowptr, const int* colind, const void* vals, const void* x, void* y,
                      const int* rows, long long nrows, long long nnz, int is_complex, const double* alpha,
                      const double* beta, int nsm, cudaStream_t s);

void launch_spmv(const int* rowptr, const int* colind, const void* vals, const void* x, void* y,
                 long long nrows, long long nnz, int is_complex, const double* alpha,
                 const double* beta, int nsm, cudaStream_t s) {
  launch_spmv_rows(rowptr, colind, vals, x, y, nullptr, nrows, nnz, is_complex, alpha, beta, nsm, s);
}

// the same product over the rows of a list (nnz: entries of those rows, for the lanes-per-row choice)
void launch_spmv_rows(const int* rowptr, const int* colind, const void* vals, const void* x, void* y,
                      const int* rows, long long nrows, long long nnz, int is_complex, const double* alpha,
                      const double* beta, int nsm, cudaStream_t s) {
  if (nrows == 0) return;
  double avg = (double)nnz / (double)nrows;
  int tpr = avg >= 48 ? 32 : avg >= 24 ? 16 : avg >= 12 ? 8 : 4;
  long long rows_per_block = 256 / tpr;
  long long grid = (nrows + rows_per_block - 1) / rows_per_block;
  long long cap = (long long)nsm * 8 * 8;
  if (grid > cap) grid = cap;
#define VLFS_SPMV_CASE(T)                                                                       \
  if (is_complex)                                                                               \
    csr_spmv_cplx<T><<<(unsigned)grid, 256, 0, s>>>(rowptr, colind, (const double2*)vals,       \
        (const double2*)x, (double2*)y, nrows, make_double2(alpha[0], alpha[1]),                \
        make_double2(beta[0], beta[1]), rows);                                                        \
  else                                                                                          \
    csr_spmv_real<T><<<(unsigned)grid, 256, 0, s>>>(rowptr, colind, (const double*)vals,        \
        (const double*)x, (double*)y, nrows, alpha[0], beta[0], rows);
  switch (tpr) {
    case 32: VLFS_SPMV_CASE(32) break;
    case 16: VLFS_SPMV_CASE(16) break;
    case 8: VLFS_SPMV_CASE(8) break;
    default: VLFS_SPMV_CASE(4) break;
  }
#undef VLFS_SPMV_CASE
  CUDA_TRY(cudaGetLastError());
}
