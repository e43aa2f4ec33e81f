// ilu0.cu - zero-fill incomplete LU of the rank's owned block as a Krylov preconditioner (sm_100a, FP64).
//
// north_star (5): "a preconditioned Krylov solve (CG / GMRES / BiCGStab with Jacobi or block-ILU0)".  The
// block is the square matrix of the rows AND columns this context owns (n = owned rows; ghost columns
// >= n are dropped: block Jacobi over the GPUs, ILU(0) inside a block).  With block_size_hint = B > 0 the
// owned block is cut further into diagonal blocks of B consecutive rows (block Jacobi of ILU(0) blocks): the
// dependency chains end at the block boundaries, and ONE launch runs every block on its own CTA.  ILU(0) in the IKJ order on the
// CSR pattern:   for i: for k < i in row i (ascending): a_ik /= a_kk;  a_ij -= a_ik a_kj  for j > k in
// row k that row i also holds.  Row i needs the finished rows k of its strict lower pattern only, so
// rows are grouped into dependency LEVELS (host analysis, once per pattern) and a level is processed
// with one warp per row: lanes split the upper part of row k and find j in row i by binary search on
// the sorted column indices; the k loop of a row stays sequential, so every a_ij receives its updates
// in ascending k exactly as the sequential algorithm does (bitwise repeatable, no atomics).
// The triangular solves z = U^-1 L^-1 r use the same levels (forward) and the levels of the transposed
// dependency (backward), one warp per row with a shuffle reduction.  Runs of small levels (<= 32 rows:
// the long dependency chains of banded matrices) are executed by ONE CTA per launch that walks the levels
// with __syncthreads between them instead of one launch per level.
//
// Replaces nothing in the serial reference (LUSolver() is exact); Jacobi- and ILU(0)-Krylov stall on the
// indefinite frequency-domain operators and converge on the Newmark matrices.
#include "ilu0.cuh"
#include <algorithm>
#include <cmath>
#include <chrono>

namespace {

template <bool CX> struct Num;
template <> struct Num<false> {
  using T = double;
  __device__ static T mul(T a, T b) { return a * b; }
  __device__ static T sub(T a, T b) { return a - b; }
  __device__ static T add(T a, T b) { return a + b; }
  __device__ static T div(T a, T b) { return a / b; }
  __device__ static T zero() { return 0.0; }
  __device__ static double abs2(T a) { return a * a; }
  __device__ static bool finite(T a) { return isfinite(a); }
  __device__ static T shfl_xor(T a, int o) { return __shfl_xor_sync(0xffffffffu, a, o); }
  __device__ static T real(double v) { return v; }
};
template <> struct Num<true> {
  using T = double2;
  __device__ static T mul(T a, T b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
  __device__ static T sub(T a, T b) { return make_double2(a.x - b.x, a.y - b.y); }
  __device__ static T add(T a, T b) { return make_double2(a.x + b.x, a.y + b.y); }
  __device__ static T div(T a, T b) {
    const double d = 1.0 / (b.x * b.x + b.y * b.y);
    return make_double2((a.x * b.x + a.y * b.y) * d, (a.y * b.x - a.x * b.y) * d);
  }
  __device__ static T zero() { return make_double2(0.0, 0.0); }
  __device__ static double abs2(T a) { return a.x * a.x + a.y * a.y; }
  __device__ static bool finite(T a) { return isfinite(a.x) && isfinite(a.y); }
  __device__ static T shfl_xor(T a, int o) {
    return make_double2(__shfl_xor_sync(0xffffffffu, a.x, o), __shfl_xor_sync(0xffffffffu, a.y, o));
  }
  __device__ static T real(double v) { return make_double2(v, 0.0); }
};

template <bool CX>
__device__ typename Num<CX>::T warp_sum(typename Num<CX>::T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = Num<CX>::add(v, Num<CX>::shfl_xor(v, o));
  return v;      // xor butterfly: every lane holds the same sum, summed in a fixed order
}

// position of column j in the sorted index range [lo, hi), or -1
__device__ __forceinline__ int find_col(const int* __restrict__ colind, int lo, int hi, int j) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int c = colind[mid];
    if (c == j) return mid;
    if (c < j) lo = mid + 1; else hi = mid;
  }
  return -1;
}

// ---- one row of the factorisation (whole warp) --------------------------------------------------------
template <bool CX>
__device__ void factor_row(int i, int r0, int r1, const int* __restrict__ rowptr, const int* __restrict__ colind,
                           const int* __restrict__ diag, typename Num<CX>::T* lu, double tiny2, int* info) {
  using N = Num<CX>;
  using T = typename N::T;
  const int lane = threadIdx.x & 31;
  const int lo = rowptr[i], hi = rowptr[i + 1], d = diag[i];
  for (int p = lo; p < d; ++p) {
    const int k = colind[p];
    if (k < r0) continue;                                  // coupling to another diagonal block: dropped (k is warp-uniform)
    const int dk = diag[k];
    const T lik = N::div(lu[p], lu[dk]);
    const int ke = rowptr[k + 1];
    for (int q = dk + 1 + lane; q < ke; q += 32) {
      const int j = colind[q];
      if (j >= r1) break;                                  // other blocks / ghost columns close a sorted row
      const int pos = find_col(colind, p + 1, hi, j);
      if (pos >= 0) lu[pos] = N::sub(lu[pos], N::mul(lik, lu[q]));
    }
    __syncwarp();
    if (lane == 0) lu[p] = lik;
    __syncwarp();
  }
  if (lane == 0) {
    const T pv = lu[d];
    if (!N::finite(pv)) atomicAdd(info + 1, 1);            // (integer counts: order-independent)
    else if (N::abs2(pv) <= tiny2) { atomicAdd(info, 1); lu[d] = N::real(sqrt(tiny2) > 0 ? sqrt(tiny2) : 1e-300); }
  }
}

template <bool CX>
__device__ void lower_row(int i, int r0, const int* __restrict__ rowptr, const int* __restrict__ colind,
                          const int* __restrict__ diag, const typename Num<CX>::T* lu,
                          const typename Num<CX>::T* r, typename Num<CX>::T* z) {
  using N = Num<CX>;
  const int lane = threadIdx.x & 31;
  typename N::T s = N::zero();
  for (int p = rowptr[i] + lane; p < diag[i]; p += 32) {
    const int k = colind[p];
    if (k >= r0) s = N::add(s, N::mul(lu[p], z[k]));
  }
  s = warp_sum<CX>(s);
  if (lane == 0) z[i] = N::sub(r[i], s);
}

template <bool CX>
__device__ void upper_row(int i, int r1, const int* __restrict__ rowptr, const int* __restrict__ colind,
                          const int* __restrict__ diag, const typename Num<CX>::T* lu, typename Num<CX>::T* z) {
  using N = Num<CX>;
  const int lane = threadIdx.x & 31;
  typename N::T s = N::zero();
  const int d = diag[i];
  for (int p = d + 1 + lane; p < rowptr[i + 1]; p += 32) {
    const int j = colind[p];
    if (j < r1) s = N::add(s, N::mul(lu[p], z[j]));
  }
  s = warp_sum<CX>(s);
  if (lane == 0) z[i] = N::div(N::sub(z[i], s), lu[d]);
}

enum { ST_FACTOR = 0, ST_LOWER = 1, ST_UPPER = 2 };

template <bool CX, int STAGE>
__device__ __forceinline__ void do_row(int i, int r0, int r1, const int* rowptr, const int* colind, const int* diag,
                                       typename Num<CX>::T* lu, const typename Num<CX>::T* r, typename Num<CX>::T* z,
                                       double tiny2, int* info) {
  if (STAGE == ST_FACTOR) factor_row<CX>(i, r0, r1, rowptr, colind, diag, lu, tiny2, info);
  else if (STAGE == ST_LOWER) lower_row<CX>(i, r0, rowptr, colind, diag, lu, r, z);
  else upper_row<CX>(i, r1, rowptr, colind, diag, lu, z);
}

// one (wide) level: one warp per row, any number of CTAs
template <bool CX, int STAGE>
__global__ void __launch_bounds__(256) ilu0_level(const int* __restrict__ rows, int nrows, int n,
                                                  const int* __restrict__ rowptr, const int* __restrict__ colind,
                                                  const int* __restrict__ diag, typename Num<CX>::T* lu,
                                                  const typename Num<CX>::T* r, typename Num<CX>::T* z, double tiny2, int* info) {
  const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (w >= nrows) return;
  do_row<CX, STAGE>(rows[w], 0, n, rowptr, colind, diag, lu, r, z, tiny2, info);
}

// a run of small levels [l0, l1) (each <= 32 rows): one CTA of 32 warps walks them
template <bool CX, int STAGE>
__global__ void __launch_bounds__(1024) ilu0_chain(const int* __restrict__ rows, const int* __restrict__ lptr, int l0, int l1,
                                                   int n, const int* __restrict__ rowptr, const int* __restrict__ colind,
                                                   const int* __restrict__ diag, typename Num<CX>::T* lu,
                                                   const typename Num<CX>::T* r, typename Num<CX>::T* z, double tiny2, int* info) {
  const int w = threadIdx.x >> 5;
  for (int l = l0; l < l1; ++l) {
    const int b = lptr[l], e = lptr[l + 1];
    if (b + w < e) do_row<CX, STAGE>(rows[b + w], 0, n, rowptr, colind, diag, lu, r, z, tiny2, info);
    __syncthreads();                 // the next level reads what this one wrote (same CTA: ordered by the barrier)
  }
}

// block Jacobi of ILU(0) blocks (block_rows > 0): diagonal block b = rows and columns [b B, (b+1) B); one CTA per
// block walks the block's own levels (rows of a block only depend on rows of the same block), all blocks at once
template <bool CX, int STAGE>
__global__ void __launch_bounds__(512) ilu0_blocks(const int* __restrict__ rows, const int* __restrict__ lptr,
                                                   const int* __restrict__ blk_lev, int block_rows, int n,
                                                   const int* __restrict__ rowptr, const int* __restrict__ colind,
                                                   const int* __restrict__ diag, typename Num<CX>::T* lu,
                                                   const typename Num<CX>::T* r, typename Num<CX>::T* z, double tiny2, int* info) {
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int r0 = blockIdx.x * block_rows;
  const int r1 = min(n, r0 + block_rows);
  const int l0 = blk_lev[blockIdx.x], l1 = blk_lev[blockIdx.x + 1] - 1;      // (one terminating entry per block)
  for (int l = l0; l < l1; ++l) {
    const int b = lptr[l], e = lptr[l + 1];
    for (int q = b + w; q < e; q += nw) do_row<CX, STAGE>(rows[q], r0, r1, rowptr, colind, diag, lu, r, z, tiny2, info);
    __syncthreads();
  }
}

template <bool CX>
__global__ void max_abs2_diag(const int* __restrict__ diag, const typename Num<CX>::T* __restrict__ v, int n, double* out) {
  // largest |a_ii|^2 (scale of the tiny-pivot threshold); one CTA, fixed order
  __shared__ double sm[256];
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) m = fmax(m, Num<CX>::abs2(v[diag[i]]));
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]); __syncthreads(); }
  if (threadIdx.x == 0) *out = sm[0];
}

template <bool CX, int STAGE>
void run_blocks(const Ilu0& P, const int* rows, const int* lptr, const int* blk_lev, const double* r, double* z,
                double tiny2, cudaStream_t s, long long* launches) {
  using T = typename Num<CX>::T;
  ilu0_blocks<CX, STAGE><<<(unsigned)P.nblocks, 512, 0, s>>>(rows, lptr, blk_lev, P.block_rows, (int)P.n, P.rowptr, P.colind, P.diag.p,
                                                            reinterpret_cast<T*>(P.lu.p), reinterpret_cast<const T*>(r),
                                                            reinterpret_cast<T*>(z), tiny2, P.info.p);
  if (launches) ++*launches;
  CUDA_TRY(cudaGetLastError());
}

template <bool CX, int STAGE>
void run_plan(const Ilu0& P, const std::vector<vlfs_ilu::Launch>& plan, const std::vector<int>& lptr_host, const int* rows,
              const int* lptr, const double* r, double* z, double tiny2, cudaStream_t s, long long* launches) {
  using T = typename Num<CX>::T;
  T* lu = reinterpret_cast<T*>(P.lu.p);
  const T* rr = reinterpret_cast<const T*>(r);
  T* zz = reinterpret_cast<T*>(z);
  for (const vlfs_ilu::Launch& L : plan) {
    if (L.chain) {
      ilu0_chain<CX, STAGE><<<1, 1024, 0, s>>>(rows, lptr, L.l0, L.l1, (int)P.n, P.rowptr, P.colind, P.diag.p, lu, rr, zz, tiny2, P.info.p);
    } else {
      const int b = lptr_host[(size_t)L.l0], nr = lptr_host[(size_t)L.l0 + 1] - b;
      ilu0_level<CX, STAGE><<<(unsigned)((nr + 7) / 8), 256, 0, s>>>(rows + b, nr, (int)P.n, P.rowptr, P.colind, P.diag.p, lu, rr, zz, tiny2, P.info.p);
    }
    if (launches) ++*launches;
  }
  CUDA_TRY(cudaGetLastError());
}

}  // namespace

// ---- host analysis (once per pattern) ---------------------------------------------------------------
std::unique_ptr<Ilu0> ilu0_symbolic(const int* rowptr_dev, const int* colind_dev, long long n, int is_complex, int block_rows,
                                    cudaStream_t s) {
  const auto t0 = std::chrono::steady_clock::now();
  std::unique_ptr<Ilu0> P(new Ilu0());
  P->n = n; P->cplx = is_complex; P->rowptr = rowptr_dev; P->colind = colind_dev; P->s = s;
  std::vector<int> rp((size_t)n + 1);
  CUDA_TRY(cudaMemcpyAsync(rp.data(), rowptr_dev, rp.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  const long long nnz = rp[(size_t)n];
  std::vector<int> ci((size_t)nnz);
  if (nnz) CUDA_TRY(cudaMemcpyAsync(ci.data(), colind_dev, ci.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  P->nnz = nnz;
  vlfs_ilu::Analysis A = vlfs_ilu::analyze(n, rp.data(), ci.data(), block_rows);      // ilu0_host.hpp (CPU-tested)
  P->block_rows = A.block_rows;
  P->nblocks = A.nblocks;
  const std::vector<int>& diag = A.diag;
  const int nl = A.lower.nlevels, nu = A.upper.nlevels;
  P->lptr_host = A.lower.lptr; P->uptr_host = A.upper.lptr;
  P->lplan = A.lower.plan; P->uplan = A.upper.plan;
  P->lrows.upload(A.lower.rows.data(), A.lower.rows.size(), s);
  P->lptr.upload(A.lower.lptr.data(), A.lower.lptr.size(), s);
  P->urows.upload(A.upper.rows.data(), A.upper.rows.size(), s);
  P->uptr.upload(A.upper.lptr.data(), A.upper.lptr.size(), s);
  if (A.block_rows) {
    P->lblk.upload(A.lower.blk.data(), A.lower.blk.size(), s);
    P->ublk.upload(A.upper.blk.data(), A.upper.blk.size(), s);
  }
  P->diag.upload(diag.data(), diag.size(), s);
  P->lu.alloc((size_t)nnz * (is_complex ? 2 : 1));
  P->info.alloc(2);
  P->scale.alloc(1);
  CUDA_TRY(cudaStreamSynchronize(s));       // the host vectors die here
  P->nlevels_lower = nl; P->nlevels_upper = nu;
  P->symbolic_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return P;
}

void ilu0_numeric(Ilu0& P, const double* vals, long long* launches) {
  const size_t bytes = (size_t)P.nnz * (P.cplx ? 2 : 1) * sizeof(double);
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  CUDA_TRY(cudaEventRecord(e0, P.s));
  if (bytes) CUDA_TRY(cudaMemcpyAsync(P.lu.p, vals, bytes, cudaMemcpyDeviceToDevice, P.s));
  P.info.zero(P.s);
  if (P.n == 0) { cudaEventDestroy(e0); cudaEventDestroy(e1); return; }
  if (P.cplx) max_abs2_diag<true><<<1, 256, 0, P.s>>>(P.diag.p, reinterpret_cast<const double2*>(P.lu.p), (int)P.n, P.scale.p);
  else max_abs2_diag<false><<<1, 256, 0, P.s>>>(P.diag.p, P.lu.p, (int)P.n, P.scale.p);
  double m2 = 0.0;
  CUDA_TRY(cudaMemcpyAsync(&m2, P.scale.p, sizeof m2, cudaMemcpyDeviceToHost, P.s));
  CUDA_TRY(cudaStreamSynchronize(P.s));
  // a pivot below eps * max|a_ii| is replaced (counted in info[0]); non-finite pivots are counted in info[1]
  const double tiny2 = m2 * 2.3e-16 * 2.3e-16;
  if (launches) ++*launches;
  if (P.block_rows) {
    if (P.cplx) run_blocks<true, ST_FACTOR>(P, P.lrows.p, P.lptr.p, P.lblk.p, nullptr, nullptr, tiny2, P.s, launches);
    else run_blocks<false, ST_FACTOR>(P, P.lrows.p, P.lptr.p, P.lblk.p, nullptr, nullptr, tiny2, P.s, launches);
  } else if (P.cplx) run_plan<true, ST_FACTOR>(P, P.lplan, P.lptr_host, P.lrows.p, P.lptr.p, nullptr, nullptr, tiny2, P.s, launches);
  else run_plan<false, ST_FACTOR>(P, P.lplan, P.lptr_host, P.lrows.p, P.lptr.p, nullptr, nullptr, tiny2, P.s, launches);
  int info[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(info, P.info.p, sizeof info, cudaMemcpyDeviceToHost, P.s));
  CUDA_TRY(cudaEventRecord(e1, P.s));
  CUDA_TRY(cudaStreamSynchronize(P.s));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  P.numeric_ms = ms;
  P.perturbed = info[0]; P.nonfinite = info[1];
}

// z = U^-1 L^-1 r   (r and z may not alias)
void ilu0_apply(Ilu0& P, const double* r, double* z, long long* launches) {
  if (P.n == 0) return;
  if (P.block_rows) {
    if (P.cplx) {
      run_blocks<true, ST_LOWER>(P, P.lrows.p, P.lptr.p, P.lblk.p, r, z, 0.0, P.s, launches);
      run_blocks<true, ST_UPPER>(P, P.urows.p, P.uptr.p, P.ublk.p, r, z, 0.0, P.s, launches);
    } else {
      run_blocks<false, ST_LOWER>(P, P.lrows.p, P.lptr.p, P.lblk.p, r, z, 0.0, P.s, launches);
      run_blocks<false, ST_UPPER>(P, P.urows.p, P.uptr.p, P.ublk.p, r, z, 0.0, P.s, launches);
    }
    return;
  }
  if (P.cplx) {
    run_plan<true, ST_LOWER>(P, P.lplan, P.lptr_host, P.lrows.p, P.lptr.p, r, z, 0.0, P.s, launches);
    run_plan<true, ST_UPPER>(P, P.uplan, P.uptr_host, P.urows.p, P.uptr.p, r, z, 0.0, P.s, launches);
  } else {
    run_plan<false, ST_LOWER>(P, P.lplan, P.lptr_host, P.lrows.p, P.lptr.p, r, z, 0.0, P.s, launches);
    run_plan<false, ST_UPPER>(P, P.uplan, P.uptr_host, P.urows.p, P.uptr.p, r, z, 0.0, P.s, launches);
  }
}
