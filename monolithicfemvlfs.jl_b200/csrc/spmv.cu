// spmv.cu - CSR sparse matrix-vector products, real and complex FP64 (sm_100a).
//
// Replaces the products Gridap/SparseArrays perform inside the Newmark step
// (src/Khabakhpasheva_time_domain.jl:258-266, src/Periodic_Beam.jl:107-116) and
// inside every Krylov iteration of the solve that stands in for `LUSolver()`
// (src/Yago_freq_domain.jl:171).  HBM-bound: nnz*(sT+4) + 4(N+1) + 2*N*sT bytes.
// A group of TPR lanes owns a row (8 lanes for the 30-100 entries per row of these operators); values are streamed, x is
// gathered through the read-only path (it lives in the 126 MB L2), the row sum is a
// shuffle reduction.
#include "common.cuh"
#include <cstdlib>

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


// ---- residual in twice the working precision ---------------------------------------------------------
// r = b - A (xh + xl), every product split exactly with an FMA (p = a*x, e = fma(a, x, -p)) and the row
// sum carried as an unevaluated pair (s, c) with TwoSum (Ogita-Rump-Oishi Dot2): the result is as
// accurate as if computed in ~106-bit arithmetic and rounded once.  With it the refinement steps of
// the LU-preconditioned solver converge to the solution of the linear system to working precision
// instead of stopping at cond(A)*eps - what an LU solution refined in extended precision gives, and
// what the 1e-8 solution gate is measured against.  HBM traffic is that of the plain product plus the
// gathers of xl; the extra flops are free next to it.
struct dd { double s, c; };
__device__ __forceinline__ void dd_add_prod(dd& a, double v, double xh, double xl) {
  const double p = __dmul_rn(v, xh);                 // (never contracted into the sum below: TwoSum needs fl(s + p))
  const double e = fma(v, xh, -p) + v * xl;
  const double t = __dadd_rn(a.s, p);
  const double bp = t - a.s;
  a.c += ((a.s - (t - bp)) + (p - bp)) + e;
  a.s = t;
}
__device__ __forceinline__ void dd_merge(dd& a, double s2, double c2) {
  const double t = a.s + s2;
  const double bp = t - a.s;
  a.c += ((a.s - (t - bp)) + (s2 - bp)) + c2;
  a.s = t;
}
__device__ __forceinline__ double dd_residual(double b, const dd& a) {      // b - (s + c), rounded once
  const double t = b - a.s;
  const double bp = t - b;
  const double err = (b - (t - bp)) + (-a.s - bp);
  return t + (err - a.c);
}

template <int TPR, bool CPLX>
__global__ void __launch_bounds__(256)
csr_residual_dd(const int* __restrict__ rowptr, const int* __restrict__ colind, const double* __restrict__ vals,
                const double* __restrict__ xh, const double* __restrict__ xl, const double* __restrict__ b,
                double* __restrict__ r, long long nrows) {
  const int lane = threadIdx.x & (TPR - 1);
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / TPR;
  const long long nrg = ((long long)gridDim.x * blockDim.x) / TPR;
  for (long long row = gid; row < nrows; row += nrg) {
    const int p0 = rowptr[row], p1 = rowptr[row + 1];
    dd re{0.0, 0.0}, im{0.0, 0.0};
    for (int p = p0 + lane; p < p1; p += TPR) {
      const int c = ld_stream(colind + p);
      if (CPLX) {
        const double2 v = ld_stream(reinterpret_cast<const double2*>(vals) + p);
        const double2 h = __ldg(reinterpret_cast<const double2*>(xh) + c), l = __ldg(reinterpret_cast<const double2*>(xl) + c);
        dd_add_prod(re, v.x, h.x, l.x); dd_add_prod(re, -v.y, h.y, l.y);
        dd_add_prod(im, v.x, h.y, l.y); dd_add_prod(im, v.y, h.x, l.x);
      } else {
        dd_add_prod(re, ld_stream(vals + p), __ldg(xh + c), __ldg(xl + c));
      }
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) {
      const double s2 = __shfl_down_sync(0xffffffffu, re.s, o, TPR), c2 = __shfl_down_sync(0xffffffffu, re.c, o, TPR);
      dd_merge(re, s2, c2);
      if (CPLX) {
        const double s3 = __shfl_down_sync(0xffffffffu, im.s, o, TPR), c3 = __shfl_down_sync(0xffffffffu, im.c, o, TPR);
        dd_merge(im, s3, c3);
      }
    }
    if (lane == 0) {
      if (CPLX) {
        const double2 bb = reinterpret_cast<const double2*>(b)[row];
        reinterpret_cast<double2*>(r)[row] = make_double2(dd_residual(bb.x, re), dd_residual(bb.y, im));
      } else {
        r[row] = dd_residual(b[row], re);
      }
    }
  }
}

// (xh, xl) += d as an unevaluated pair, renormalised
template <bool CPLX>
__global__ void dd_axpy(double* __restrict__ xh, double* __restrict__ xl, const double* __restrict__ d, long long n) {
  const long long len = CPLX ? 2 * n : n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x) {
    const double a = xh[i], dv = d[i];
    const double t = a + dv;
    const double bp = t - a;
    const double lo = ((a - (t - bp)) + (dv - bp)) + xl[i];
    const double h = t + lo;
    xh[i] = h;
    xl[i] = lo - (h - t);
  }
}

}  // namespace

void launch_residual_dd(const int* rowptr, const int* colind, const void* vals, const double* xh, const double* xl,
                        const double* b, double* r, long long nrows, long long nnz, int is_complex, int nsm, cudaStream_t s) {
  if (nrows == 0) return;
  const double avg = (double)nnz / (double)nrows;
  const int tpr = avg >= 128 ? 32 : 8;      // (same finding as for the plain product: few lanes per row)
  long long grid = (nrows + 256 / tpr - 1) / (256 / tpr);
  if (grid > (long long)nsm * 64) grid = (long long)nsm * 64;
  const double* v = (const double*)vals;
  if (is_complex) {
    if (tpr == 32) csr_residual_dd<32, true><<<(unsigned)grid, 256, 0, s>>>(rowptr, colind, v, xh, xl, b, r, nrows);
    else csr_residual_dd<8, true><<<(unsigned)grid, 256, 0, s>>>(rowptr, colind, v, xh, xl, b, r, nrows);
  } else {
    if (tpr == 32) csr_residual_dd<32, false><<<(unsigned)grid, 256, 0, s>>>(rowptr, colind, v, xh, xl, b, r, nrows);
    else csr_residual_dd<8, false><<<(unsigned)grid, 256, 0, s>>>(rowptr, colind, v, xh, xl, b, r, nrows);
  }
  CUDA_TRY(cudaGetLastError());
}

void launch_dd_axpy(double* xh, double* xl, const double* d, long long n, int is_complex, int nsm, cudaStream_t s) {
  if (n == 0) return;
  long long grid = (n + 255) / 256;
  if (grid > (long long)nsm * 16) grid = (long long)nsm * 16;
  if (is_complex) dd_axpy<true><<<(unsigned)grid, 256, 0, s>>>(xh, xl, d, n);
  else dd_axpy<false><<<(unsigned)grid, 256, 0, s>>>(xh, xl, d, n);
  CUDA_TRY(cudaGetLastError());
}

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
  // lanes per row: about 8-10 entries per lane.  Measured on Y1 (76 entries per row, B200, profiles/r2_spmv_lanes.md):
  // 32 lanes 0.79 / 0.64 of the HBM peak (complex / real), 16 lanes 0.94 / 0.78, 8 lanes 0.95 / 0.85, 4 lanes 0.87 / 0.77 -
  // with 32 lanes a 76-entry row leaves a 12-entry tail on a quarter of the warp and pays a 5-step reduction per row
  int tpr = avg >= 256 ? 32 : avg >= 128 ? 16 : avg >= 20 ? 8 : 4;
  // tuning hooks (lanes per row of the real / complex product)
  static const int tpr_real = getenv("VLFS_SPMV_TPR_REAL") ? atoi(getenv("VLFS_SPMV_TPR_REAL")) : 0;
  static const int tpr_cplx = getenv("VLFS_SPMV_TPR_CPLX") ? atoi(getenv("VLFS_SPMV_TPR_CPLX")) : 0;
  const int forced = is_complex ? tpr_cplx : tpr_real;
  if (forced == 4 || forced == 8 || forced == 16 || forced == 32) tpr = forced;
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
