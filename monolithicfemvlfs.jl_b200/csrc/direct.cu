// direct.cu - GPU multifrontal sparse LU (FP64 real / complex, sm_100a).
//
// This is the B200 replacement of `LUSolver()` = UMFPACK `lu(A)` + `ldiv!` behind
// `solve(op)` (src/Yago_freq_domain.jl:171, src/Khabakhpasheva_freq_domain.jl:241,
// src/Multi_geo_freq_domain.jl:135) and behind `Newmark(LUSolver(),...)`
// (src/Khabakhpasheva_time_domain.jl:258-259, src/Periodic_Beam.jl:107-108).  The
// monolithic matrices are indefinite wave operators on which Jacobi/ILU0 Krylov stalls
// (solver study), so the factorisation itself runs on the GPU:
//   symbolic (host threads, once per pattern; symbolic_host.hpp): geometric nested dissection on
//     the DOF coordinates, separator tree -> supernodes, front index sets, level schedule;
//   numeric (device): dense fronts in one pool; per tree level, batched over fronts:
//     extend-add of the children's Schur complements, then a blocked right-looking partial LU
//     (32-column panels inside outer blocks of 128 pivots: diagonal-block LU in shared memory,
//     panel triangular solves, strip updates per panel, trailing update once per outer block
//     with 64x64 register-tiled FP64 tiles);
//   solve (device): level-scheduled forward / backward sweeps over the fronts, one launch per
//     panel; unknowns with role "keep" stay a Schur boundary (substructuring across GPUs).
// Pivoting: the elimination order of the fronts is static; inside every 32-row pivot block rows are
// interchanged by threshold partial pivoting (diag_lu), pivots that stay tiny are perturbed and
// counted, non-finite pivots are counted (direct_pivot_status -> VLFS_PIVOT_PERTURBED /
// VLFS_ERR_SINGULAR at the C ABI).  The Krylov driver uses the factorisation as a right
// preconditioner, so the reported residual is always the true one and refinement recovers the
// accuracy a perturbed pivot costs.
#include "common.cuh"
#include "direct.cuh"
#include "symbolic_host.hpp"
#include <algorithm>
#include <chrono>
#include <numeric>
#include <cmath>

namespace {

constexpr int NB = 32;      // panel width
constexpr int TS = 64;      // Schur tile

template <bool C> struct Num;
template <> struct Num<false> {
  using T = double;
  __device__ static T zero() { return 0.0; }
  __device__ static T mul(T a, T b) { return a * b; }
  __device__ static T fnma(T a, T b, T c) { return fma(-a, b, c); }      // c - a*b
  __device__ static T div(T a, T b) { return a / b; }
  __device__ static T add(T a, T b) { return a + b; }
  __device__ static T shfl(T a, int src) { return __shfl_sync(0xffffffffu, a, src); }
  __device__ static T shfl_down(T a, int o) { return __shfl_down_sync(0xffffffffu, a, o); }
};
template <> struct Num<true> {
  using T = double2;
  __device__ static T zero() { return make_double2(0, 0); }
  __device__ static T mul(T a, T b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
  __device__ static T fnma(T a, T b, T c) {
    c.x = fma(-a.x, b.x, c.x); c.x = fma(a.y, b.y, c.x);
    c.y = fma(-a.x, b.y, c.y); c.y = fma(-a.y, b.x, c.y);
    return c;
  }
  __device__ static T div(T a, T b) {
    double d = 1.0 / (b.x * b.x + b.y * b.y);
    return make_double2((a.x * b.x + a.y * b.y) * d, (a.y * b.x - a.x * b.y) * d);
  }
  __device__ static T add(T a, T b) { return make_double2(a.x + b.x, a.y + b.y); }
  __device__ static T shfl(T a, int src) {
    return make_double2(__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src));
  }
  __device__ static T shfl_down(T a, int o) {
    return make_double2(__shfl_down_sync(0xffffffffu, a.x, o), __shfl_down_sync(0xffffffffu, a.y, o));
  }
};

struct FrontMeta {
  long long off;      // offset of the nf x nf row-major front in the pool (scalars)
  long long woff;     // offset of the front's work vector (nf scalars)
  int first;          // first pivot position (permuted numbering)
  int np, nf;
  int upd_ptr;        // into upd_idx (nf - np entries, ascending permuted positions)
  int child[2];       // child fronts or -1
  int cmap_ptr;       // into cmap: local index in the parent of each update index
  int parent;
};

__device__ __forceinline__ int local_index(const FrontMeta& F, const int* __restrict__ upd_idx, int p) {
  if (p < F.first + F.np) return p - F.first;
  const int* u = upd_idx + F.upd_ptr;
  int lo = 0, hi = F.nf - F.np;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (u[mid] < p) lo = mid + 1; else hi = mid;
  }
  return F.np + lo;
}

template <bool C>
__global__ void scatter_matrix(const int* __restrict__ rowind, const int* __restrict__ colind,
                               const typename Num<C>::T* __restrict__ vals, long long nnz,
                               const int* __restrict__ pos, int n_elim, const int* __restrict__ front_of_pos,
                               const FrontMeta* __restrict__ meta, const int* __restrict__ upd_idx,
                               typename Num<C>::T* __restrict__ pool) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < nnz;
       p += (long long)gridDim.x * blockDim.x) {
    int pi = pos[rowind[p]], pj = pos[colind[p]];
    if (pi < 0 || pj < 0) continue;        // ignored row / column (ghosts of a distributed system)
    if (pi >= n_elim && pj >= n_elim) continue;   // kept x kept: stays outside the factorisation
    int f = front_of_pos[pi < pj ? pi : pj];
    const FrontMeta F = meta[f];
    int li = local_index(F, upd_idx, pi), lj = local_index(F, upd_idx, pj);
    pool[F.off + (long long)li * F.nf + lj] = vals[p];
  }
}

__global__ void build_cmap(const FrontMeta* __restrict__ meta, int nfronts, const int* __restrict__ upd_idx,
                           int* __restrict__ cmap) {
  int f = blockIdx.x;
  if (f >= nfronts) return;
  const FrontMeta F = meta[f];
  if (F.parent < 0) return;
  const FrontMeta P = meta[F.parent];
  const int* u = upd_idx + F.upd_ptr;
  for (int k = threadIdx.x; k < F.nf - F.np; k += blockDim.x)
    cmap[F.cmap_ptr + k] = local_index(P, upd_idx, u[k]);
}

// parent += Schur complement of child `slot`; one launch per slot keeps the sum order fixed.
// A CTA of 64 x 4 threads covers 16 rows: four independent read-modify-write chains per thread.
template <bool C>
__global__ void __launch_bounds__(256)
extend_add(const FrontMeta* __restrict__ meta, const int* __restrict__ list, int slot,
           const int* __restrict__ cmap, typename Num<C>::T* __restrict__ pool) {
  using N = Num<C>;
  const FrontMeta P = meta[list[blockIdx.z]];
  const int c = P.child[slot];
  if (c < 0) return;
  const FrontMeta Cf = meta[c];
  const int nu = Cf.nf - Cf.np;
  const int a0 = blockIdx.y * 16 + threadIdx.y;
  if (blockIdx.y * 16 >= nu) return;
  const int* cm = cmap + Cf.cmap_ptr;
  long long prow[4], crow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + 4 * i;
    prow[i] = a < nu ? P.off + (long long)cm[a] * P.nf : -1;
    crow[i] = Cf.off + (long long)(Cf.np + a) * Cf.nf + Cf.np;
  }
  for (int b = blockIdx.x * 64 + threadIdx.x; b < nu; b += gridDim.x * 64) {
    const int cb = cm[b];
    typename N::T v[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (prow[i] >= 0) { v[i] = pool[crow[i] + b]; d[i] = pool[prow[i] + cb]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (prow[i] >= 0) pool[prow[i] + cb] = N::add(d[i], v[i]);
  }
}

// |re| + |im| (LAPACK cabs1): cheap magnitude for pivot comparisons
__device__ __forceinline__ double abs1(double a) { return fabs(a); }
__device__ __forceinline__ double abs1(double2 a) { return fabs(a.x) + fabs(a.y); }
__device__ __forceinline__ bool finite1(double a) { return isfinite(a); }
__device__ __forceinline__ bool finite1(double2 a) { return isfinite(a.x) && isfinite(a.y); }
__device__ __forceinline__ double perturbed(double a, double tau) { return a < 0.0 ? -tau : tau; }
__device__ __forceinline__ double2 perturbed(double2 a, double tau) {
  const double m = fabs(a.x) + fabs(a.y);
  return m > 0.0 ? make_double2(a.x / m * tau, a.y / m * tau) : make_double2(tau, 0.0);
}

// LU of the diagonal block of panel k (in place) with THRESHOLD PARTIAL PIVOTING INSIDE THE BLOCK:
// the diagonal entry stays the pivot while |a_jj| >= 0.1 max_r |a_rj|, otherwise the largest entry
// of the column inside the block is swapped in (UMFPACK's strategy, restricted to the 32 rows of
// the block so that the front structure is static).  perm[first + k0 + r] = row of the block that
// ends up at position r; panel_trsm (U part) and the forward sweep apply it to the rows k0..k1 of
// the front / the work vector.  A pivot that is still tiny against the block (< 2^-43 of its
// largest entry) or not finite is counted in info[0] / info[1]; tiny pivots are replaced by
// +-tau (static perturbation, as SuperLU_DIST does) so that no Inf/NaN spreads silently.
template <bool C>
__global__ void __launch_bounds__(NB * NB)
diag_lu(const FrontMeta* __restrict__ meta, const int* __restrict__ list, int k,
        typename Num<C>::T* __restrict__ pool, int* __restrict__ perm, int* __restrict__ info) {
  using N = Num<C>;
  __shared__ typename N::T sm[NB][NB + 1];
  __shared__ double smax[NB];
  __shared__ int s_p, s_perm[NB];
  const FrontMeta F = meta[list[blockIdx.x]];
  const int k0 = k * NB;
  if (k0 >= F.np) return;
  const int kb = min(NB, F.np - k0);
  const int r = threadIdx.y, c = threadIdx.x;      // warp = row r
  typename N::T* D = pool + F.off + (long long)k0 * F.nf + k0;
  double mine = 0.0;
  if (r < kb && c < kb) { sm[r][c] = D[(long long)r * F.nf + c]; mine = abs1(sm[r][c]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine = fmax(mine, __shfl_xor_sync(0xffffffffu, mine, o));
  if (c == 0) { smax[r] = mine; s_perm[r] = r; }
  __syncthreads();
  double scale = 0.0;
  for (int q = 0; q < NB; ++q) scale = fmax(scale, smax[q]);
  const double tau = scale * 1.1368683772161603e-13;       // 2^-43
  for (int j = 0; j < kb; ++j) {
    if (r == 0) {                                           // warp 0: pivot search in column j
      double a = (c >= j && c < kb) ? abs1(sm[c][j]) : -1.0;
      const double ajj = __shfl_sync(0xffffffffu, a, j);
      int p = c;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double a2 = __shfl_xor_sync(0xffffffffu, a, o);
        const int p2 = __shfl_xor_sync(0xffffffffu, p, o);
        if (a2 > a || (a2 == a && p2 < p)) { a = a2; p = p2; }
      }
      if (c == 0) s_p = (ajj >= 0.1 * a) ? j : p;           // NaNs compare false: keeps the search result
    }
    __syncthreads();
    const int p = s_p;
    if (r == j) {                                           // warp j: swap rows j and p of the block
      if (p != j && c < kb) { typename N::T t = sm[j][c]; sm[j][c] = sm[p][c]; sm[p][c] = t; }
      if (c == j) {
        if (p != j) { const int t = s_perm[j]; s_perm[j] = s_perm[p]; s_perm[p] = t; }
        const typename N::T pv = sm[j][j];
        if (!finite1(pv)) atomicAdd(info + 1, 1);
        else if (abs1(pv) < tau || abs1(pv) == 0.0) {
          sm[j][j] = perturbed(pv, tau > 0.0 ? tau : 1.0);
          atomicAdd(info, 1);
        }
      }
    }
    __syncthreads();
    if (c == j && r > j && r < kb) sm[r][j] = N::div(sm[r][j], sm[j][j]);
    __syncthreads();
    if (r > j && c > j && r < kb && c < kb) sm[r][c] = N::fnma(sm[r][j], sm[j][c], sm[r][c]);
    __syncthreads();
  }
  if (r < kb && c < kb) D[(long long)r * F.nf + c] = sm[r][c];
  if (r == 0 && c < kb) perm[F.first + k0 + c] = s_perm[c];
}

// L-panel: rows below the diagonal block times U11^{-1};  U-panel: L11^{-1} times the columns
// to the right.  blockIdx.y = 0: L-panel (thread per row), 1: U-panel (thread per column).
template <bool C>
__global__ void __launch_bounds__(128)
panel_trsm(const FrontMeta* __restrict__ meta, const int* __restrict__ list, int k,
           typename Num<C>::T* __restrict__ pool, const int* __restrict__ perm) {
  using N = Num<C>;
  __shared__ typename N::T sD[NB][NB + 1];
  __shared__ int sperm[NB];
  const FrontMeta F = meta[list[blockIdx.z]];
  const int k0 = k * NB;
  if (k0 >= F.np) return;
  const int kb = min(NB, F.np - k0);
  const int k1 = k0 + kb;
  const int idx = k1 + blockIdx.x * 128 + threadIdx.x;
  if (k1 + blockIdx.x * 128 >= F.nf) return;
  typename N::T* Fp = pool + F.off;
  for (int t = threadIdx.x; t < kb * kb; t += 128)
    sD[t / kb][t % kb] = Fp[(long long)(k0 + t / kb) * F.nf + k0 + t % kb];
  if (threadIdx.x < NB) sperm[threadIdx.x] = threadIdx.x < kb ? perm[F.first + k0 + threadIdx.x] : 0;
  __syncthreads();
  if (idx >= F.nf) return;
  typename N::T x[NB];
  if (blockIdx.y == 0) {
    typename N::T* row = Fp + (long long)idx * F.nf + k0;
#pragma unroll
    for (int c = 0; c < NB; ++c) x[c] = c < kb ? row[c] : N::zero();
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      if (c < kb) {
        typename N::T v = x[c];
#pragma unroll
        for (int r = 0; r < c; ++r) v = N::fnma(x[r], sD[r][c], v);
        x[c] = N::div(v, sD[c][c]);
      }
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) if (c < kb) row[c] = x[c];
  } else {
    typename N::T* col = Fp + (long long)k0 * F.nf + idx;
#pragma unroll
    for (int r = 0; r < NB; ++r) x[r] = r < kb ? col[(long long)sperm[r] * F.nf] : N::zero();   // rows as pivoted
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      if (r < kb) {
        typename N::T v = x[r];
#pragma unroll
        for (int q = 0; q < r; ++q) v = N::fnma(sD[r][q], x[q], v);
        x[r] = v;
      }
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) if (r < kb) col[(long long)r * F.nf] = x[r];
  }
}

// Blocked right-looking update  F[i][j] -= sum_{r in [d0,d1)} F[i][r] F[r][j]  over one region of
// the front.  The pivots advance in outer blocks of `outer` panels; inside an outer block ending
// at Kend only the block's own column strip (mode 1: i >= d1, d1 <= j < Kend) and row strip
// (mode 2: d1 <= i < Kend, j >= Kend) are updated after every 32-wide panel, with depth 32; the
// trailing matrix (mode 0: i, j >= Kend) is visited ONCE per outer block with the whole block as
// depth, so its tiles are read and written `outer` times less often and a CTA amortises its
// prologue over `outer` x 32 contraction steps.
template <bool C>
__global__ void __launch_bounds__(256)
schur_update(const FrontMeta* __restrict__ meta, const int* __restrict__ list, int kf, int nk, int mode,
             int kend_panel, typename Num<C>::T* __restrict__ pool) {
  using N = Num<C>;
  extern __shared__ double smraw[];
  typename N::T(*sA)[NB + 1] = reinterpret_cast<typename N::T(*)[NB + 1]>(smraw);
  typename N::T(*sB)[TS] = reinterpret_cast<typename N::T(*)[TS]>(sA + TS);
  const FrontMeta F = meta[list[blockIdx.z]];
  const int d0 = kf * NB;
  if (d0 >= F.np) return;
  const int d1 = min((kf + nk) * NB, F.np);
  const int Kend = min(kend_panel * NB, F.np);
  int i0, j0, ilim, jlim;
  if (mode == 0) { i0 = Kend + blockIdx.y * TS; j0 = Kend + blockIdx.x * TS; ilim = F.nf; jlim = F.nf; }
  else if (mode == 1) { i0 = d1 + blockIdx.y * TS; j0 = d1 + blockIdx.x * TS; ilim = F.nf; jlim = Kend; }
  else { i0 = d1 + blockIdx.y * TS; j0 = Kend + blockIdx.x * TS; ilim = Kend; jlim = F.nf; }
  if (i0 >= ilim || j0 >= jlim) return;
  typename N::T* Fp = pool + F.off;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  typename N::T acc[4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[m][n] = N::zero();
  for (int kb0 = d0; kb0 < d1; kb0 += NB) {
    const int kb = min(NB, d1 - kb0);
    if (kb0 > d0) __syncthreads();
    for (int t = tid; t < TS * NB; t += 256) {
      int r = t / NB, c = t % NB;
      int i = i0 + r;
      sA[r][c] = (i < ilim && c < kb) ? Fp[(long long)i * F.nf + kb0 + c] : N::zero();
    }
    for (int t = tid; t < NB * TS; t += 256) {
      int r = t / TS, c = t % TS;
      int j = j0 + c;
      sB[r][c] = (j < jlim && r < kb) ? Fp[(long long)(kb0 + r) * F.nf + j] : N::zero();
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < NB; ++kk) {
      typename N::T a[4], b[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) a[m] = sA[ty + 16 * m][kk];
#pragma unroll
      for (int n = 0; n < 4; ++n) b[n] = sB[kk][tx + 16 * n];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[m][n] = N::fnma(a[m], b[n], acc[m][n]);
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    int i = i0 + ty + 16 * m;
    if (i >= ilim) continue;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      int j = j0 + tx + 16 * n;
      if (j < jlim) {
        typename N::T* p = Fp + (long long)i * F.nf + j;
        *p = N::add(*p, acc[m][n]);
      }
    }
  }
}

// The same update on the FP64 tensor cores (mma.sync.m8n8k4.f64; measured on this B200: 37.1 TFLOP/s
// against 33.6 of the FMA pipe, scripts/ubench/dmma_peak.cu) with the operand tiles brought in by
// cp.async into a double-buffered shared-memory ring.  The register-tiled version above loads one
// shared-memory operand per 2 real FMAs and waits for its global loads between two barriers (ncu:
// long-scoreboard 14 cycles per issue, 17 % issue slots); here a warp-level 8x8x4 product needs 2
// operand loads per thread for 16 FMAs and the loads of stage s+1 fly during the products of stage s.
// Tile 64 x 64 per CTA, 8 warps as 2 x 4, a warp owns 32 x 16 = 4 x 2 mma tiles, depth 16 per stage.
// Row strides of 20 (A) and 66 / 68 (B, complex / real) scalars keep the fragment loads of a quarter /
// half warp in distinct banks.  The accumulators collect +A*B and are subtracted at the end.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? BYTES : 0;                   // src-size 0: the destination is zero filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(d), "l"(gsrc), "n"(BYTES), "r"(n) : "memory");
}
constexpr int MMA_KD = 16;                           // depth of a stage
constexpr int MMA_LDA = MMA_KD + 4;
template <bool C> struct MmaLd { static constexpr int B = C ? TS + 2 : TS + 4; };
template <bool C> constexpr size_t mma_smem() { return 2 * (size_t)(TS * MMA_LDA + MMA_KD * MmaLd<C>::B) * sizeof(typename Num<C>::T); }

template <bool C>
__global__ void __launch_bounds__(256, 2)       // (3 CTAs/SM fit the shared memory but spill at 80 registers: 221 vs 215 ms)
schur_update_mma(const FrontMeta* __restrict__ meta, const int* __restrict__ list, int kf, int nk, int mode,
                 int kend_panel, typename Num<C>::T* __restrict__ pool) {
  using N = Num<C>;
  using T = typename N::T;
  constexpr int P = C ? 2 : 1;
  constexpr int LDB = MmaLd<C>::B;
  constexpr int STAGE = TS * MMA_LDA + MMA_KD * LDB;           // scalars per stage
  extern __shared__ __align__(16) double smraw[];
  T* sm = reinterpret_cast<T*>(smraw);
  const FrontMeta F = meta[list[blockIdx.z]];
  const int d0 = kf * NB;
  if (d0 >= F.np) return;
  const int d1 = min((kf + nk) * NB, F.np);
  const int Kend = min(kend_panel * NB, F.np);
  int i0, j0, ilim, jlim;
  if (mode == 0) { i0 = Kend + blockIdx.y * TS; j0 = Kend + blockIdx.x * TS; ilim = F.nf; jlim = F.nf; }
  else if (mode == 1) { i0 = d1 + blockIdx.y * TS; j0 = d1 + blockIdx.x * TS; ilim = F.nf; jlim = Kend; }
  else { i0 = d1 + blockIdx.y * TS; j0 = Kend + blockIdx.x * TS; ilim = Kend; jlim = F.nf; }
  if (i0 >= ilim || j0 >= jlim) return;
  T* Fp = pool + F.off;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp >> 2, wc = warp & 3;           // warp tile: rows wr*32.., columns wc*16..
  const int g = lane >> 2, t4 = lane & 3;            // groupID, threadID_in_group of the fragment layouts
  double acc[4][2][P][2];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int p = 0; p < P; ++p) acc[m][n][p][0] = acc[m][n][p][1] = 0.0;
  const int nstage = (d1 - d0 + MMA_KD - 1) / MMA_KD;
  auto issue = [&](int st) {
    const int kb0 = d0 + st * MMA_KD, kb = min(MMA_KD, d1 - kb0);
    T* sA = sm + (size_t)(st & 1) * STAGE;
    T* sB = sA + TS * MMA_LDA;
#pragma unroll
    for (int t = tid; t < TS * MMA_KD; t += 256) {               // A: rows i0.., columns kb0..
      const int r = t / MMA_KD, c = t % MMA_KD;
      const bool ok = i0 + r < ilim && c < kb;
      cp_async_zfill<sizeof(T)>(sA + r * MMA_LDA + c, ok ? Fp + (long long)(i0 + r) * F.nf + kb0 + c : Fp, ok);
    }
#pragma unroll
    for (int t = tid; t < MMA_KD * TS; t += 256) {               // B: rows kb0.., columns j0..
      const int r = t / TS, c = t % TS;
      const bool ok = j0 + c < jlim && r < kb;
      cp_async_zfill<sizeof(T)>(sB + r * LDB + c, ok ? Fp + (long long)(kb0 + r) * F.nf + j0 + c : Fp, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
  for (int st = 0; st < nstage; ++st) {
    if (st + 1 < nstage) { issue(st + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const T* sA = sm + (size_t)(st & 1) * STAGE;
    const T* sB = sA + TS * MMA_LDA;
#pragma unroll
    for (int k4 = 0; k4 < MMA_KD / 4; ++k4) {
      T a[4], b[2];
#pragma unroll
      for (int m = 0; m < 4; ++m) a[m] = sA[(wr * 32 + m * 8 + g) * MMA_LDA + k4 * 4 + t4];
#pragma unroll
      for (int n = 0; n < 2; ++n) b[n] = sB[(k4 * 4 + t4) * LDB + wc * 16 + n * 8 + g];
      // (the two products into the same accumulator are issued 16 DMMAs apart: ncu showed the pipe
      //  throttling on the dependent back-to-back pairs)
      if constexpr (C) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int n = 0; n < 2; ++n) { dmma884(acc[m][n][0], a[m].x, b[n].x); dmma884(acc[m][n][1], a[m].x, b[n].y); }
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int n = 0; n < 2; ++n) { dmma884(acc[m][n][0], -a[m].y, b[n].y); dmma884(acc[m][n][1], a[m].y, b[n].x); }
      } else {
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int n = 0; n < 2; ++n) dmma884(acc[m][n][0], a[m], b[n]);
      }
    }
    __syncthreads();                                   // the next issue overwrites this buffer
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int i = i0 + wr * 32 + m * 8 + g;
    if (i >= ilim) continue;
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const int j = j0 + wc * 16 + n * 8 + 2 * t4;
      T* p = Fp + (long long)i * F.nf + j;
      if constexpr (C) {
        if (j < jlim) p[0] = make_double2(p[0].x - acc[m][n][0][0], p[0].y - acc[m][n][1][0]);
        if (j + 1 < jlim) p[1] = make_double2(p[1].x - acc[m][n][0][1], p[1].y - acc[m][n][1][1]);
      } else {
        if (j < jlim) p[0] -= acc[m][n][0][0];
        if (j + 1 < jlim) p[1] -= acc[m][n][0][1];
      }
    }
  }
}

// ---- solve kernels --------------------------------------------------------------------
// b (original numbering, n_total entries) -> bp (eliminated positions only)
template <bool C>
__global__ void permute_in(const typename Num<C>::T* __restrict__ b, const int* __restrict__ pos,
                           typename Num<C>::T* __restrict__ bp, long long n_total, int n_elim) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x) {
    const int p = pos[i];
    if (p >= 0 && p < n_elim) bp[p] = b[i];
  }
}
// xp (eliminated + kept positions) -> x (original numbering); ignored entries are left alone
template <bool C>
__global__ void permute_out(const typename Num<C>::T* __restrict__ xp, const int* __restrict__ pos,
                            typename Num<C>::T* __restrict__ x, long long n_total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x) {
    const int p = pos[i];
    if (p >= 0) x[i] = xp[p];
  }
}
// S[keep a][keep b] += trailing block of a root front (its Schur complement on the kept set)
template <bool C>
__global__ void add_root_schur(FrontMeta F, const int* __restrict__ upd_idx, int n_elim, int n_keep,
                               const typename Num<C>::T* __restrict__ pool, typename Num<C>::T* __restrict__ S) {
  using N = Num<C>;
  const int nu = F.nf - F.np;
  const int* u = upd_idx + F.upd_ptr;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)nu * nu;
       t += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(t / nu), b = (int)(t - (long long)a * nu);
    typename N::T* dst = S + (size_t)(u[a] - n_elim) * n_keep + (u[b] - n_elim);
    *dst = N::add(*dst, pool[F.off + (long long)(F.np + a) * F.nf + F.np + b]);
  }
}
// front of a block-tridiagonal chain <- dense interface matrix S (ng x ng): pivots [o, o+np), update
// [o+np, o+nf); the update x update corner belongs to the next front
template <bool C>
__global__ void chain_fill(FrontMeta F, const typename Num<C>::T* __restrict__ S, long long ng,
                           typename Num<C>::T* __restrict__ pool) {
  using N = Num<C>;
  const long long total = (long long)F.nf * F.nf;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / F.nf), j = (int)(t - (long long)i * F.nf);
    pool[F.off + t] = (i >= F.np && j >= F.np) ? N::zero() : S[(long long)(F.first + i) * ng + F.first + j];
  }
}

// g[keep a] += update part of a root's work vector (forward sweep);  xp[kept] = xk (backward)
template <bool C>
__global__ void add_root_rhs(FrontMeta F, const int* __restrict__ upd_idx, int n_elim,
                             const typename Num<C>::T* __restrict__ w, typename Num<C>::T* __restrict__ g) {
  using N = Num<C>;
  const int nu = F.nf - F.np;
  const int* u = upd_idx + F.upd_ptr;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < nu; a += gridDim.x * blockDim.x)
    g[u[a] - n_elim] = N::add(g[u[a] - n_elim], w[F.woff + F.np + a]);
}

// w = [b_P ; 0] + contributions of the children (fixed order: slot 0 then slot 1)
template <bool C>
__global__ void fwd_gather(const FrontMeta* __restrict__ meta, const int* __restrict__ list,
                           const int* __restrict__ cmap, const typename Num<C>::T* __restrict__ bp,
                           typename Num<C>::T* __restrict__ w) {
  using N = Num<C>;
  const FrontMeta F = meta[list[blockIdx.x]];
  typename N::T* wf = w + F.woff;
  for (int i = threadIdx.x; i < F.nf; i += blockDim.x) wf[i] = i < F.np ? bp[F.first + i] : N::zero();
  __syncthreads();
  for (int s = 0; s < 2; ++s) {
    const int c = F.child[s];
    if (c < 0) continue;
    const FrontMeta Cf = meta[c];
    const typename N::T* wc = w + Cf.woff + Cf.np;
    const int* cm = cmap + Cf.cmap_ptr;
    for (int a = threadIdx.x; a < Cf.nf - Cf.np; a += blockDim.x) wf[cm[a]] = N::add(wf[cm[a]], wc[a]);
    __syncthreads();
  }
}

// forward panel step: solve the unit-lower diagonal block, then update all rows below.  Every
// block redoes the tiny triangular solve (one launch per panel); the diagonal block is staged in
// shared memory first so that the 32-step shuffle chain carries no global-memory latency.
// Block 0 stores the solved pivots y into `bp` (whose pivot part the front's gather has already
// consumed): nothing in this launch reads it, so no second "commit" launch is needed.
template <bool C>
__global__ void __launch_bounds__(256)
fwd_panel(const FrontMeta* __restrict__ meta, const int* __restrict__ list, int k,
          const typename Num<C>::T* __restrict__ pool, typename Num<C>::T* __restrict__ w,
          typename Num<C>::T* __restrict__ bp, const int* __restrict__ perm) {
  using N = Num<C>;
  __shared__ typename N::T sx[NB];
  __shared__ typename N::T sD[NB][NB + 1];
  const FrontMeta F = meta[list[blockIdx.y]];
  const int k0 = k * NB;
  if (k0 >= F.np) return;
  const int kb = min(NB, F.np - k0);
  const int k1 = k0 + kb;
  if (blockIdx.x > 0 && k1 + blockIdx.x * 256 >= F.nf) return;
  const typename N::T* Fp = pool + F.off;
  typename N::T* wf = w + F.woff;
  for (int t = threadIdx.x; t < kb * kb; t += 256) {
    const int r = t / kb, c = t - r * kb;
    sD[r][c] = Fp[(long long)(k0 + r) * F.nf + k0 + c];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    typename N::T x = r < kb ? wf[k0 + perm[F.first + k0 + r]] : N::zero();   // row interchanges of the block
    for (int c = 0; c < kb; ++c) {
      typename N::T xc = N::shfl(x, c);
      if (r > c && r < kb) x = N::fnma(sD[r][c], xc, x);
    }
    if (r < kb) {
      sx[r] = x;
      if (blockIdx.x == 0) bp[F.first + k0 + r] = x;
    }
  }
  __syncthreads();
  const int i = k1 + blockIdx.x * 256 + threadIdx.x;
  if (i < F.nf) {
    const typename N::T* row = Fp + (long long)i * F.nf + k0;
    typename N::T v = wf[i];
    for (int c = 0; c < kb; ++c) v = N::fnma(row[c], sx[c], v);
    wf[i] = v;
  }
}

// backward: r_P = y_P - U12 x_U  (warp per pivot row), x_U gathered from the solved vector, y_P
// from the forward sweep (stored in the permuted right-hand side)
template <bool C>
__global__ void __launch_bounds__(256)
bwd_gemv(const FrontMeta* __restrict__ meta, const int* __restrict__ list, const int* __restrict__ upd_idx,
         const typename Num<C>::T* __restrict__ pool, const typename Num<C>::T* __restrict__ xp,
         const typename Num<C>::T* __restrict__ y, typename Num<C>::T* __restrict__ w) {
  using N = Num<C>;
  const FrontMeta F = meta[list[blockIdx.y]];
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= F.np) return;
  const int nu = F.nf - F.np;
  const typename N::T* Fr = pool + F.off + (long long)row * F.nf + F.np;
  const int* u = upd_idx + F.upd_ptr;
  typename N::T acc = N::zero();
  for (int a = lane; a < nu; a += 32) acc = N::fnma(Fr[a], xp[u[a]], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = N::add(acc, N::shfl_down(acc, o));
  if (lane == 0) w[F.woff + row] = N::add(y[F.first + row], acc);
}

// backward panel step (descending k): solve the upper diagonal block, update the rows above;
// block 0 stores the solution of the panel into the permuted solution vector
template <bool C>
__global__ void __launch_bounds__(256)
bwd_panel(const FrontMeta* __restrict__ meta, const int* __restrict__ list, int k,
          const typename Num<C>::T* __restrict__ pool, typename Num<C>::T* __restrict__ w,
          typename Num<C>::T* __restrict__ xp) {
  using N = Num<C>;
  __shared__ typename N::T sx[NB];
  __shared__ typename N::T sD[NB][NB + 1];
  const FrontMeta F = meta[list[blockIdx.y]];
  const int k0 = k * NB;
  if (k0 >= F.np) return;
  if (blockIdx.x > 0 && blockIdx.x * 256 >= k0) return;
  const int kb = min(NB, F.np - k0);
  const typename N::T* Fp = pool + F.off;
  typename N::T* wf = w + F.woff;
  for (int t = threadIdx.x; t < kb * kb; t += 256) {
    const int r = t / kb, c = t - r * kb;
    sD[r][c] = Fp[(long long)(k0 + r) * F.nf + k0 + c];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    typename N::T x = r < kb ? wf[k0 + r] : N::zero();
    for (int c = kb - 1; c >= 0; --c) {
      if (r == c) x = N::div(x, sD[c][c]);
      typename N::T xc = N::shfl(x, c);
      if (r < c) x = N::fnma(sD[r][c], xc, x);
    }
    if (r < kb) {
      sx[r] = x;
      if (blockIdx.x == 0) xp[F.first + k0 + r] = x;
    }
  }
  __syncthreads();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < k0) {
    const typename N::T* row = Fp + (long long)i * F.nf + k0;
    typename N::T v = wf[i];
    for (int c = 0; c < kb; ++c) v = N::fnma(row[c], sx[c], v);
    wf[i] = v;
  }
}

}  // namespace

// =======================================================================================
// host side: symbolic analysis
// =======================================================================================
struct DirectSolver {
  long long n = 0, nnz = 0;          // n: eliminated unknowns
  long long n_total = 0, n_keep = 0; // original vector length; kept (Schur boundary) unknowns
  std::vector<int> roots;
  bool dense = false;                // single dense front (interface system)
  int cplx = 0, nsm = 148;
  cudaStream_t s = nullptr;
  long long* launches = nullptr;
  int nfronts = 0, nlevels = 0, maxnf = 0;
  long long pool_elems = 0, w_elems = 0, fill = 0;
  double flops = 0;
  std::vector<FrontMeta> meta;
  std::vector<std::vector<int>> level_lists;
  std::vector<int> level_maxnp, level_maxnf, level_maxnu_child;
  DevBuf<FrontMeta> d_meta;
  DevBuf<int> d_pos, d_front_of_pos, d_upd, d_cmap, d_lists;
  std::vector<int> list_off;
  DevBuf<double> pool, w, bp, xp;
  DevBuf<int> d_perm;                // row interchanges inside the 32-row pivot blocks (diag_lu)
  DevBuf<int> d_info;                // {perturbed pivots, non-finite pivots} of the last factorisation
  long long n_perturbed = 0, n_nonfinite = 0;
  double symbolic_ms = 0, numeric_ms = 0;
  void alloc_pivot_state() {
    d_perm.alloc((size_t)std::max<long long>(n, 1));
    d_info.alloc(2);
  }
};

namespace {

template <typename F>
float timed(cudaStream_t s, F&& f) {
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  CUDA_TRY(cudaEventRecord(e0, s));
  f();
  CUDA_TRY(cudaEventRecord(e1, s));
  CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return ms;
}

}  // namespace

void DirectDeleter::operator()(DirectSolver* p) const { delete p; }

DirectPtr direct_symbolic(const int* d_rowptr, const int* d_colind, long long n_total,
                          const signed char* role, long long nnz, const double* coords, int dim, int leaf,
                          int is_complex, int nsm, cudaStream_t s, long long* launches) {
  // role[i] (host, n_total entries): 0 = eliminate, 1 = keep as Schur boundary (never a pivot; the
  // root fronts end with the Schur complement on these unknowns), 2 = ignore (ghosts).
  DirectPtr D(new DirectSolver());
  D->n_total = n_total;
  D->nnz = nnz; D->cplx = is_complex; D->nsm = nsm; D->s = s; D->launches = launches;
  if (leaf < 8) leaf = 96;
  auto t0 = std::chrono::steady_clock::now();
  auto tlast = t0;
  const bool dbg = getenv("VLFS_DEBUG") != nullptr;
  auto lap = [&](const char* what) {
    if (!dbg) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[vlfs] symbolic: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - tlast).count());
    tlast = t;
  };
  std::vector<int> rp(n_total + 1), ci(nnz);
  CUDA_TRY(cudaMemcpy(rp.data(), d_rowptr, (n_total + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(ci.data(), d_colind, nnz * sizeof(int), cudaMemcpyDeviceToHost));
  lap("pattern to host");
  // host analysis (symbolic_host.hpp): A + A^T, geometric nested dissection, update sets
  vlfs_sym::Analysis R;
  vlfs_sym::analyse(n_total, rp.data(), ci.data(), role, coords, dim, leaf, R);
  if (dbg)
    fprintf(stderr, "[vlfs] symbolic: symmetrise %.1f ms, nested dissection %.1f ms, update sets %.1f ms (%u threads)\n",
            R.ms_sym, R.ms_nd, R.ms_upd, vlfs_sym::host_threads());
  const long long n = R.n;
  D->n = n; D->n_keep = R.n_keep;
  const int nf_ = (int)R.nodes.size();
  D->nfronts = nf_;
  const std::vector<int>& pos = R.pos;
  const std::vector<int>& front_of_pos = R.front_of_pos;
  const std::vector<int>& first = R.first;
  const std::vector<int>& parent = R.parent;
  const std::vector<int>& level = R.level;
  const std::vector<std::vector<int>>& upd = R.upd;
  struct { const std::vector<vlfs_sym::TreeNode>& nodes; } nd{R.nodes};
  std::vector<int>().swap(rp); std::vector<int>().swap(ci);
  lap("update sets");
  // meta
  D->meta.resize(nf_);
  std::vector<int> upd_flat;
  long long off = 0, woff = 0;
  int cm = 0;
  D->nlevels = 0;
  for (int f = 0; f < nf_; ++f) {
    FrontMeta& M = D->meta[f];
    M.first = first[f]; M.np = (int)nd.nodes[f].piv.size(); M.nf = M.np + (int)upd[f].size();
    M.off = off; M.woff = woff; M.upd_ptr = (int)upd_flat.size(); M.cmap_ptr = cm; M.parent = parent[f];
    M.child[0] = nd.nodes[f].child[0]; M.child[1] = nd.nodes[f].child[1];
    off += (long long)M.nf * M.nf; woff += M.nf; cm += M.nf - M.np;
    upd_flat.insert(upd_flat.end(), upd[f].begin(), upd[f].end());
    D->maxnf = std::max(D->maxnf, M.nf);
    D->nlevels = std::max(D->nlevels, level[f] + 1);
    const double np = M.np, nu = M.nf - M.np;
    D->fill += (long long)(np * np + 2 * np * nu);
    D->flops += (2.0 / 3.0) * np * np * np + 2 * np * np * nu + 2 * np * nu * nu;
  }
  D->pool_elems = off; D->w_elems = woff;
  D->level_lists.assign(D->nlevels, {});
  for (int f = 0; f < nf_; ++f) D->level_lists[level[f]].push_back(f);
  {
    // the launches are batched over the fronts of a level through gridDim.y / gridDim.z (<= 65 535):
    // a longer level is processed as several consecutive ones (its fronts are independent)
    int zmax = 65535;
    if (const char* e = getenv("VLFS_LU_ZMAX")) { const int v = atoi(e); if (v >= 1 && v < zmax) zmax = v; }
    std::vector<std::vector<int>> split;
    for (const std::vector<int>& lst : D->level_lists)
      for (size_t o = 0; o < lst.size(); o += (size_t)zmax)
        split.emplace_back(lst.begin() + o, lst.begin() + std::min(lst.size(), o + (size_t)zmax));
    D->level_lists.swap(split);
    D->nlevels = (int)D->level_lists.size();
  }
  D->level_maxnp.assign(D->nlevels, 0); D->level_maxnf.assign(D->nlevels, 0);
  D->level_maxnu_child.assign(D->nlevels, 0);
  std::vector<int> lists_flat;
  D->list_off.clear();
  for (int L = 0; L < D->nlevels; ++L) {
    D->list_off.push_back((int)lists_flat.size());
    for (int f : D->level_lists[L]) {
      lists_flat.push_back(f);
      D->level_maxnp[L] = std::max(D->level_maxnp[L], D->meta[f].np);
      D->level_maxnf[L] = std::max(D->level_maxnf[L], D->meta[f].nf);
      for (int c : D->meta[f].child)
        if (c >= 0) D->level_maxnu_child[L] = std::max(D->level_maxnu_child[L], D->meta[c].nf - D->meta[c].np);
    }
  }
  lap("front metadata");
  // upload
  D->d_meta.upload(D->meta.data(), nf_, s);
  D->d_pos.upload(pos.data(), n_total, s);
  D->d_front_of_pos.upload(front_of_pos.data(), n, s);
  if (upd_flat.empty()) upd_flat.push_back(0);
  D->d_upd.upload(upd_flat.data(), upd_flat.size(), s);
  D->d_cmap.alloc(std::max(cm, 1));
  D->d_lists.upload(lists_flat.data(), lists_flat.size(), s);
  build_cmap<<<nf_, 128, 0, s>>>(D->d_meta.p, nf_, D->d_upd.p, D->d_cmap.p);
  ++*launches;
  const size_t sc = is_complex ? 2 : 1;
  {
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const double need = (double)D->pool_elems * sc * sizeof(double);
    if (need > 0.92 * (double)free_b) {
      char buf[256];
      snprintf(buf, sizeof buf, "sparse LU needs %.1f GB for its fronts (largest %d), %.1f GB free",
               need / 1e9, D->maxnf, free_b / 1e9);
      throw std::string(buf);
    }
  }
  D->pool.alloc((size_t)D->pool_elems * sc);
  D->w.alloc((size_t)D->w_elems * sc);
  D->bp.alloc((size_t)(n + D->n_keep) * sc);
  D->xp.alloc((size_t)(n + D->n_keep) * sc);
  D->xp.zero(s);
  D->alloc_pivot_state();
  for (int f = 0; f < nf_; ++f) if (parent[f] < 0) D->roots.push_back(f);
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaGetLastError());
  lap("upload + pool allocation");
  D->symbolic_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return D;
}

template <bool C>
static void factor_levels(DirectSolver& D) {
  using T = typename Num<C>::T;
  cudaStream_t s = D.s;
  T* pool = reinterpret_cast<T*>(D.pool.p);
  static const bool use_mma = !(getenv("VLFS_LU_MMA") && atoi(getenv("VLFS_LU_MMA")) == 0);
  const size_t smem = use_mma ? mma_smem<C>() : (size_t)(TS * (NB + 1) + NB * TS) * sizeof(T);
  auto schur = use_mma ? schur_update_mma<C> : schur_update<C>;
  CUDA_TRY(cudaFuncSetAttribute(schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static const int outer_env = getenv("VLFS_LU_OUTER") ? atoi(getenv("VLFS_LU_OUTER")) : 4;
  const int outer = outer_env < 1 ? 1 : outer_env;   // panels per outer block (4 = 128 pivots)
  D.d_info.zero(s);
  for (int L = 0; L < D.nlevels; ++L) {
    const int nfr = (int)D.level_lists[L].size();
    const int* list = D.d_lists.p + D.list_off[L];
    if (D.level_maxnu_child[L] > 0) {
      const int nu = D.level_maxnu_child[L];
      dim3 grid(std::min((nu + 63) / 64, 8), (nu + 15) / 16, nfr), block(64, 4);
      for (int slot = 0; slot < 2; ++slot) {
        extend_add<C><<<grid, block, 0, s>>>(D.d_meta.p, list, slot, D.d_cmap.p, pool);
        ++*D.launches;
      }
    }
    const int npan = (D.level_maxnp[L] + NB - 1) / NB;
    const int maxnf = D.level_maxnf[L];
    for (int K0 = 0; K0 < npan; K0 += outer) {
      const int K1 = std::min(K0 + outer, npan);          // outer block = panels [K0, K1)
      for (int k = K0; k < K1; ++k) {
        diag_lu<C><<<nfr, dim3(NB, NB), 0, s>>>(D.d_meta.p, list, k, pool, D.d_perm.p, D.d_info.p);
        ++*D.launches;
        const int rem = maxnf - k * NB;                   // upper bound of the rows/columns from panel k on
        if (rem <= 0) continue;
        dim3 g2((rem + 127) / 128, 2, nfr);
        panel_trsm<C><<<g2, 128, 0, s>>>(D.d_meta.p, list, k, pool, D.d_perm.p);
        ++*D.launches;
        const int strip = (K1 - k - 1) * NB;              // rest of the outer block after panel k
        if (strip > 0) {
          dim3 gc((strip + TS - 1) / TS, (rem + TS - 1) / TS, nfr);
          schur<<<gc, 256, smem, s>>>(D.d_meta.p, list, k, 1, 1, K1, pool);
          dim3 gr((rem + TS - 1) / TS, (strip + TS - 1) / TS, nfr);
          schur<<<gr, 256, smem, s>>>(D.d_meta.p, list, k, 1, 2, K1, pool);
          *D.launches += 2;
        }
      }
      const int rem = maxnf - K0 * NB;                    // >= nf - Kend of every front still active
      if (rem > 0) {
        dim3 g3((rem + TS - 1) / TS, (rem + TS - 1) / TS, nfr);
        schur<<<g3, 256, smem, s>>>(D.d_meta.p, list, K0, K1 - K0, 0, K1, pool);
        ++*D.launches;
      }
    }
  }
  CUDA_TRY(cudaGetLastError());
}

// pivot statistics of the factorisation just enqueued (synchronises the stream)
static void read_pivot_info(DirectSolver& D) {
  int h[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(h, D.d_info.p, sizeof h, cudaMemcpyDeviceToHost, D.s));
  CUDA_TRY(cudaStreamSynchronize(D.s));
  D.n_perturbed = h[0]; D.n_nonfinite = h[1];
}

template <bool C>
static void numeric_impl(DirectSolver& D, const int* rowind, const int* colind, const double* vals) {
  using T = typename Num<C>::T;
  cudaStream_t s = D.s;
  T* pool = reinterpret_cast<T*>(D.pool.p);
  D.pool.zero(s);
  {
    long long g = (D.nnz + 255) / 256;
    long long cap = (long long)D.nsm * 16;
    scatter_matrix<C><<<(unsigned)std::min(g, cap), 256, 0, s>>>(rowind, colind, reinterpret_cast<const T*>(vals), D.nnz,
        D.d_pos.p, (int)D.n, D.d_front_of_pos.p, D.d_meta.p, D.d_upd.p, pool);
    ++*D.launches;
  }
  factor_levels<C>(D);
}

void direct_numeric(DirectSolver& D, const int* rowind, const int* colind, const double* vals) {
  D.numeric_ms = timed(D.s, [&]() {
    if (D.cplx) numeric_impl<true>(D, rowind, colind, vals);
    else numeric_impl<false>(D, rowind, colind, vals);
  });
  read_pivot_info(D);
}

void direct_pivot_status(const DirectSolver& D, long long* perturbed, long long* nonfinite) {
  if (perturbed) *perturbed = D.n_perturbed;
  if (nonfinite) *nonfinite = D.n_nonfinite;
}

// forward sweep: y = L^-1 b on the eliminated unknowns; g_keep (optional, accumulated into) receives
// the update parts of the roots = -A_KE A_EE^-1 b_E
template <bool C>
static void forward_impl(DirectSolver& D, const double* b, double* g_keep) {
  using T = typename Num<C>::T;
  cudaStream_t s = D.s;
  const T* pool = reinterpret_cast<const T*>(D.pool.p);
  T* w = reinterpret_cast<T*>(D.w.p);
  T* bp = reinterpret_cast<T*>(D.bp.p);
  const unsigned vg = (unsigned)std::min<long long>((D.n_total + 255) / 256, (long long)D.nsm * 8);
  permute_in<C><<<vg, 256, 0, s>>>(reinterpret_cast<const T*>(b), D.d_pos.p, bp, D.n_total, (int)D.n);
  ++*D.launches;
  for (int L = 0; L < D.nlevels; ++L) {
    const int nfr = (int)D.level_lists[L].size();
    const int* list = D.d_lists.p + D.list_off[L];
    fwd_gather<C><<<nfr, 256, 0, s>>>(D.d_meta.p, list, D.d_cmap.p, bp, w);
    ++*D.launches;
    const int npan = (D.level_maxnp[L] + NB - 1) / NB;
    for (int k = 0; k < npan; ++k) {
      const int rem = D.level_maxnf[L] - k * NB;
      dim3 g((std::max(rem, 1) + 255) / 256, nfr);
      fwd_panel<C><<<g, 256, 0, s>>>(D.d_meta.p, list, k, pool, w, bp, D.d_perm.p);
      ++*D.launches;
    }
  }
  if (g_keep && D.n_keep > 0)
    for (int f : D.roots) {
      const FrontMeta& F = D.meta[f];
      if (F.nf == F.np) continue;
      add_root_rhs<C><<<(F.nf - F.np + 255) / 256, 256, 0, s>>>(F, D.d_upd.p, (int)D.n, w, reinterpret_cast<T*>(g_keep));
      ++*D.launches;
    }
  CUDA_TRY(cudaGetLastError());
}

// backward sweep with known kept values x_keep (optional; zero when absent); x in original numbering
template <bool C>
static void backward_impl(DirectSolver& D, const double* x_keep, double* x) {
  using T = typename Num<C>::T;
  cudaStream_t s = D.s;
  const T* pool = reinterpret_cast<const T*>(D.pool.p);
  T* w = reinterpret_cast<T*>(D.w.p);
  T* xp = reinterpret_cast<T*>(D.xp.p);
  if (D.n_keep > 0) {
    if (x_keep) CUDA_TRY(cudaMemcpyAsync(xp + D.n, x_keep, (size_t)D.n_keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
    else CUDA_TRY(cudaMemsetAsync(xp + D.n, 0, (size_t)D.n_keep * sizeof(T), s));
  }
  for (int L = D.nlevels - 1; L >= 0; --L) {
    const int nfr = (int)D.level_lists[L].size();
    const int* list = D.d_lists.p + D.list_off[L];
    dim3 gg((D.level_maxnp[L] + 7) / 8, nfr);
    bwd_gemv<C><<<gg, 256, 0, s>>>(D.d_meta.p, list, D.d_upd.p, pool, xp, reinterpret_cast<const T*>(D.bp.p), w);
    ++*D.launches;
    const int npan = (D.level_maxnp[L] + NB - 1) / NB;
    for (int k = npan - 1; k >= 0; --k) {
      dim3 g((std::max(k * NB, 1) + 255) / 256, nfr);
      bwd_panel<C><<<g, 256, 0, s>>>(D.d_meta.p, list, k, pool, w, xp);
      ++*D.launches;
    }
  }
  const unsigned vg = (unsigned)std::min<long long>((D.n_total + 255) / 256, (long long)D.nsm * 8);
  permute_out<C><<<vg, 256, 0, s>>>(xp, D.d_pos.p, reinterpret_cast<T*>(x), D.n_total);
  ++*D.launches;
  CUDA_TRY(cudaGetLastError());
}

void direct_forward(DirectSolver& D, const double* b_dev, double* g_keep_dev) {
  if (D.cplx) forward_impl<true>(D, b_dev, g_keep_dev);
  else forward_impl<false>(D, b_dev, g_keep_dev);
}
void direct_backward(DirectSolver& D, const double* x_keep_dev, double* x_dev) {
  if (D.cplx) backward_impl<true>(D, x_keep_dev, x_dev);
  else backward_impl<false>(D, x_keep_dev, x_dev);
}
void direct_solve(DirectSolver& D, const double* b_dev, double* x_dev) {
  direct_forward(D, b_dev, nullptr);
  direct_backward(D, nullptr, x_dev);
}

long long direct_nkeep(const DirectSolver& D) { return D.n_keep; }

// S (n_keep x n_keep, row-major, accumulated into) += -A_KE A_EE^-1 A_EK after direct_numeric
void direct_schur(DirectSolver& D, double* S_dev) {
  for (int f : D.roots) {
    const FrontMeta& F = D.meta[f];
    const long long nu = F.nf - F.np;
    if (nu == 0) continue;
    const unsigned grid = (unsigned)std::min<long long>((nu * nu + 255) / 256, (long long)D.nsm * 16);
    if (D.cplx) add_root_schur<true><<<grid, 256, 0, D.s>>>(F, D.d_upd.p, (int)D.n, (int)D.n_keep,
                                                            reinterpret_cast<const double2*>(D.pool.p), reinterpret_cast<double2*>(S_dev));
    else add_root_schur<false><<<grid, 256, 0, D.s>>>(F, D.d_upd.p, (int)D.n, (int)D.n_keep, D.pool.p, S_dev);
    ++*D.launches;
  }
  CUDA_TRY(cudaGetLastError());
}

// ---- one dense n x n system factorised with the front kernels (interface Schur complement) ----
DirectPtr direct_dense(long long n, int is_complex, int nsm, cudaStream_t s, long long* launches) {
  DirectPtr D(new DirectSolver());
  D->n = n; D->n_total = n; D->n_keep = 0; D->nnz = 0; D->cplx = is_complex; D->nsm = nsm; D->s = s;
  D->launches = launches; D->dense = true;
  D->nfronts = 1; D->nlevels = 1; D->maxnf = (int)n;
  FrontMeta M{};
  M.off = 0; M.woff = 0; M.first = 0; M.np = (int)n; M.nf = (int)n; M.upd_ptr = 0; M.child[0] = M.child[1] = -1;
  M.cmap_ptr = 0; M.parent = -1;
  D->meta.assign(1, M);
  D->level_lists.assign(1, std::vector<int>{0});
  D->level_maxnp.assign(1, (int)n); D->level_maxnf.assign(1, (int)n); D->level_maxnu_child.assign(1, 0);
  D->list_off.assign(1, 0);
  D->roots.assign(1, 0);
  D->pool_elems = n * n; D->w_elems = n;
  D->fill = n * n; D->flops = (2.0 / 3.0) * (double)n * n * n;
  std::vector<int> pos(n), zero1(1, 0);
  std::iota(pos.begin(), pos.end(), 0);
  D->d_meta.upload(D->meta.data(), 1, s);
  D->d_pos.upload(pos.data(), n, s);
  D->d_front_of_pos.alloc(1);
  D->d_upd.upload(zero1.data(), 1, s);
  D->d_cmap.alloc(1);
  D->d_lists.upload(zero1.data(), 1, s);
  const size_t sc = is_complex ? 2 : 1;
  D->pool.alloc((size_t)n * n * sc);
  D->w.alloc((size_t)n * sc); D->bp.alloc((size_t)n * sc); D->xp.alloc((size_t)n * sc);
  D->alloc_pivot_state();
  CUDA_TRY(cudaStreamSynchronize(s));
  return D;
}

// ---- one partial front: np pivots, nf - np kept unknowns (a rank's link of the distributed interface
// chain: own cut = pivots, next cut = update set).  Same kernels; np = 0 is legal (nothing to do).
DirectPtr direct_front(long long np, long long nf, int is_complex, int nsm, cudaStream_t s, long long* launches) {
  DirectPtr D(new DirectSolver());
  D->n = np; D->n_total = nf; D->n_keep = nf - np; D->nnz = 0; D->cplx = is_complex; D->nsm = nsm; D->s = s;
  D->launches = launches; D->dense = true;
  D->nfronts = 1; D->nlevels = 1; D->maxnf = (int)nf;
  FrontMeta M{};
  M.off = 0; M.woff = 0; M.first = 0; M.np = (int)np; M.nf = (int)nf; M.upd_ptr = 0; M.child[0] = M.child[1] = -1;
  M.cmap_ptr = 0; M.parent = -1;
  D->meta.assign(1, M);
  D->level_lists.assign(1, std::vector<int>{0});
  D->level_maxnp.assign(1, (int)np); D->level_maxnf.assign(1, (int)nf); D->level_maxnu_child.assign(1, 0);
  D->list_off.assign(1, 0);
  D->roots.assign(1, 0);
  D->pool_elems = nf * nf; D->w_elems = nf;
  const double p = (double)np, u = (double)(nf - np);
  D->fill = (long long)(p * p + 2 * p * u); D->flops = (2.0 / 3.0) * p * p * p + 2 * p * p * u + 2 * p * u * u;
  std::vector<int> pos(nf), upd, zero1(1, 0);
  std::iota(pos.begin(), pos.end(), 0);
  for (long long k = np; k < nf; ++k) upd.push_back((int)k);
  if (upd.empty()) upd.push_back(0);
  D->d_meta.upload(D->meta.data(), 1, s);
  D->d_pos.upload(pos.data(), nf, s);
  D->d_front_of_pos.alloc(1);
  D->d_upd.upload(upd.data(), upd.size(), s);
  D->d_cmap.alloc(1);
  D->d_lists.upload(zero1.data(), 1, s);
  const size_t sc = is_complex ? 2 : 1;
  D->pool.alloc((size_t)nf * nf * sc);
  D->w.alloc((size_t)nf * sc); D->bp.alloc((size_t)nf * sc); D->xp.alloc((size_t)nf * sc);
  D->xp.zero(s);
  D->alloc_pivot_state();
  CUDA_TRY(cudaStreamSynchronize(s));
  return D;
}
double* direct_pool(DirectSolver& D) { return D.pool.p; }
// factorises the np pivots of the front whose nf x nf entries the caller wrote into direct_pool();
// asynchronous (no host synchronisation: the chain of ranks is ordered by the stream and NCCL)
void direct_front_factor_async(DirectSolver& D) {
  if (D.n == 0) return;
  if (D.cplx) factor_levels<true>(D); else factor_levels<false>(D);
}
void direct_front_read_status(DirectSolver& D) { read_pivot_info(D); }

// ---- block-tridiagonal interface system (slab cuts in rank order): a chain of fronts, front r has
// the unknowns of cut r as pivots and those of cut r+1 as update set - the same kernels, O(sum n_r^3)
DirectPtr direct_chain(int nblocks, const long long* sizes, int is_complex, int nsm, cudaStream_t s,
                       long long* launches) {
  DirectPtr D(new DirectSolver());
  long long n = 0;
  for (int r = 0; r < nblocks; ++r) n += sizes[r];
  D->n = n; D->n_total = n; D->n_keep = 0; D->nnz = 0; D->cplx = is_complex; D->nsm = nsm; D->s = s;
  D->launches = launches; D->dense = true;
  D->nfronts = nblocks; D->nlevels = nblocks;
  D->meta.resize(nblocks);
  std::vector<int> upd_flat, lists(nblocks);
  long long off = 0, woff = 0, o = 0;
  int cm = 0;
  D->level_lists.assign(nblocks, {});
  D->level_maxnp.assign(nblocks, 0); D->level_maxnf.assign(nblocks, 0); D->level_maxnu_child.assign(nblocks, 0);
  D->list_off.resize(nblocks);
  for (int r = 0; r < nblocks; ++r) {
    FrontMeta& M = D->meta[r];
    M.first = (int)o; M.np = (int)sizes[r]; M.nf = M.np + (r + 1 < nblocks ? (int)sizes[r + 1] : 0);
    M.off = off; M.woff = woff; M.upd_ptr = (int)upd_flat.size(); M.cmap_ptr = cm;
    M.parent = r + 1 < nblocks ? r + 1 : -1; M.child[0] = r > 0 ? r - 1 : -1; M.child[1] = -1;
    for (int k = M.np; k < M.nf; ++k) upd_flat.push_back((int)(o + k));
    off += (long long)M.nf * M.nf; woff += M.nf; cm += M.nf - M.np; o += sizes[r];
    D->maxnf = std::max(D->maxnf, M.nf);
    D->level_lists[r].push_back(r); lists[r] = r; D->list_off[r] = r;
    D->level_maxnp[r] = M.np; D->level_maxnf[r] = M.nf;
    D->level_maxnu_child[r] = r > 0 ? D->meta[r - 1].nf - D->meta[r - 1].np : 0;
    const double np = M.np, nu = M.nf - M.np;
    D->fill += (long long)(np * np + 2 * np * nu);
    D->flops += (2.0 / 3.0) * np * np * np + 2 * np * np * nu + 2 * np * nu * nu;
  }
  D->roots.assign(1, nblocks - 1);
  D->pool_elems = off; D->w_elems = woff;
  std::vector<int> pos(n);
  std::iota(pos.begin(), pos.end(), 0);
  if (upd_flat.empty()) upd_flat.push_back(0);
  D->d_meta.upload(D->meta.data(), nblocks, s);
  D->d_pos.upload(pos.data(), n, s);
  D->d_front_of_pos.alloc(1);
  D->d_upd.upload(upd_flat.data(), upd_flat.size(), s);
  D->d_cmap.alloc(std::max(cm, 1));
  D->d_lists.upload(lists.data(), lists.size(), s);
  build_cmap<<<nblocks, 128, 0, s>>>(D->d_meta.p, nblocks, D->d_upd.p, D->d_cmap.p);
  const size_t sc = is_complex ? 2 : 1;
  D->pool.alloc((size_t)off * sc);
  D->w.alloc((size_t)woff * sc); D->bp.alloc((size_t)n * sc); D->xp.alloc((size_t)n * sc);
  D->alloc_pivot_state();
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaGetLastError());
  return D;
}

void direct_chain_factor(DirectSolver& D, const double* S_dev) {
  D.numeric_ms = timed(D.s, [&]() {
    for (int r = 0; r < D.nfronts; ++r) {
      const FrontMeta& F = D.meta[r];
      const unsigned grid = (unsigned)std::min<long long>(((long long)F.nf * F.nf + 255) / 256, (long long)D.nsm * 16);
      if (D.cplx) chain_fill<true><<<grid, 256, 0, D.s>>>(F, reinterpret_cast<const double2*>(S_dev), D.n, reinterpret_cast<double2*>(D.pool.p));
      else chain_fill<false><<<grid, 256, 0, D.s>>>(F, S_dev, D.n, D.pool.p);
      ++*D.launches;
    }
    if (D.cplx) factor_levels<true>(D); else factor_levels<false>(D);
  });
  read_pivot_info(D);
}

void direct_dense_factor(DirectSolver& D, const double* S_dev) {
  const size_t sc = D.cplx ? 2 : 1;
  D.numeric_ms = timed(D.s, [&]() {
    CUDA_TRY(cudaMemcpyAsync(D.pool.p, S_dev, (size_t)D.n * D.n * sc * sizeof(double), cudaMemcpyDeviceToDevice, D.s));
    if (D.cplx) factor_levels<true>(D); else factor_levels<false>(D);
  });
  read_pivot_info(D);
}

void direct_info(const DirectSolver& D, double* out8) {
  out8[0] = D.nfronts; out8[1] = D.nlevels; out8[2] = D.maxnf; out8[3] = (double)D.fill;
  out8[4] = D.flops * (D.cplx ? 4.0 : 1.0); out8[5] = (double)D.pool_elems * (D.cplx ? 16 : 8);
  out8[6] = D.symbolic_ms; out8[7] = D.numeric_ms;
}
