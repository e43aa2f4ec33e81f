// integrate.cu - numeric phase of the assembly into the fixed CSR pattern (sm_100a, FP64
// throughout).
//
// Replaces Gridap's lazy cell-matrix evaluation and COO fill inside
// `AffineFEOperator(a,l,X,Y)` / `assemble_matrix_and_vector`
// (reference call sites: src/Yago_freq_domain.jl:161-167,
// src/Khabakhpasheva_freq_domain.jl:226-240, src/Khabakhpasheva_time_domain.jl:237-257,
// src/Periodic_Beam.jl:96-104, src/Multi_geo_freq_domain.jl:125-131).
//
// The numeric phase is three stages (api.cu::assemble_impl):
//  1. class tables: cells with equal quantised Jacobians, table variants and coefficient fields
//     share one element block.  build_class_tables tabulates the reference-type products
//     (∫∇u·∇v, value x value on constant-Jacobian cells) and integrate_cells<D> integrates the
//     remaining terms on ONE representative cell per class into the same contiguous tables.
//  2. gather_compact / gather_reference: slot-centric pass - one thread per CSR slot (two slots per
//     thread in the compact variant) sums its COO entries in COO order from the class tables and
//     writes the value once (also the zero fill); no atomics.
//  3. integrate_cells<D> on all cells for what stays cell-dependent (non-affine cells, fields
//     that differ in every cell): one CTA per cell, software-pipelined over its cells.  Geometry
//     (Jacobian pseudo-inverse, measure, normals) two cells ahead, operator tables (value /
//     gradient / directional derivative / Laplacian / Hessian / C:Hessian / normal contractions)
//     one cell ahead, 4x4 register-tiled contraction per (block, matrix, re|im) target,
//     red.global.add.f64 through the tile-major `dest`.
//  rhs_affine: linear forms on constant-Jacobian cells, warp per cell, on a side stream.
//  functional_cells<D>: ∑∫|OP(u_h) − f|² dμ of FE functions (L2 errors, energies).
#include "common.cuh"
#include <cstdlib>
#include <algorithm>
#include <cuda/barrier>

namespace {

template <int N>
__device__ __forceinline__ double det_inv(const double (&A)[3][3], double (&I)[3][3]) {
  if (N == 1) {
    I[0][0] = 1.0 / A[0][0];
    return A[0][0];
  } else if (N == 2) {
    double d = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    double r = 1.0 / d;
    I[0][0] = A[1][1] * r; I[0][1] = -A[0][1] * r;
    I[1][0] = -A[1][0] * r; I[1][1] = A[0][0] * r;
    return d;
  } else {
    double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
    double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
    double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    double d = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    double r = 1.0 / d;
    I[0][0] = c00 * r;
    I[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * r;
    I[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * r;
    I[1][0] = c01 * r;
    I[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * r;
    I[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * r;
    I[2][0] = c02 * r;
    I[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * r;
    I[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * r;
    return d;
  }
}

__device__ __forceinline__ double det_inv_n(int n, const double (&A)[3][3], double (&I)[3][3]) {
  if (n == 1) return det_inv<1>(A, I);
  if (n == 2) return det_inv<2>(A, I);
  return det_inv<3>(A, I);
}

// geometry record in shared memory per (slot, q): pinv[D][3] (9) + normal[3] = 12 doubles
constexpr int GEO_STRIDE = 12;

template <int D>
__device__ void slot_geometry(const SlotDev& S, long long cell, int q, int nq, double* rec,
                              double* meas_out, int measure_kind) {
  const int var = S.variant ? S.variant[cell] : 0;
  const double* gg = S.geo_grad + ((size_t)var * nq + q) * S.nv * S.dref;
  const double* X = S.coords + (size_t)cell * S.nv * D;
  double Jt[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int v = 0; v < S.nv; ++v)
    for (int a = 0; a < S.dref; ++a) {
      double g = gg[v * S.dref + a];
#pragma unroll
      for (int b = 0; b < D; ++b) Jt[a][b] = fma(g, X[v * D + b], Jt[a][b]);
    }
  double pinv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // [b][a]
  double meas;
  if (S.dref == D) {
    double I[3][3];
    double d = det_inv_n(D, Jt, I);
    for (int b = 0; b < D; ++b)
      for (int a = 0; a < D; ++a) pinv[b][a] = I[b][a];
    meas = fabs(d);
  } else {
    double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Gi[3][3];
    for (int a = 0; a < S.dref; ++a)
      for (int c = 0; c < S.dref; ++c) {
        double s = 0;
#pragma unroll
        for (int b = 0; b < D; ++b) s = fma(Jt[a][b], Jt[c][b], s);
        G[a][c] = s;
      }
    double d = det_inv_n(S.dref, G, Gi);
    for (int b = 0; b < D; ++b)
      for (int a = 0; a < S.dref; ++a) {
        double s = 0;
        for (int c = 0; c < S.dref; ++c) s = fma(Jt[c][b], Gi[c][a], s);
        pinv[b][a] = s;
      }
    meas = sqrt(d);
  }
  for (int b = 0; b < 3; ++b)
    for (int a = 0; a < 3; ++a) rec[b * 3 + a] = (b < D) ? pinv[b][a] : 0.0;
  double n[3] = {0, 0, 0};
  if (S.ref_normal) {
    const double* nr = S.ref_normal + var * S.dref;
    double nn = 0;
    for (int b = 0; b < D; ++b) {
      double s = 0;
      for (int a = 0; a < S.dref; ++a) s = fma(pinv[b][a], nr[a], s);
      n[b] = s;
      nn = fma(s, s, nn);
    }
    nn = 1.0 / sqrt(nn);
    for (int b = 0; b < D; ++b) n[b] *= nn;
  }
  rec[9] = n[0]; rec[10] = n[1]; rec[11] = n[2];
  if (meas_out) {
    if (measure_kind == VLFS_MEASURE_FACET) {
      if (S.ref_tangent) {
        const double* tr = S.ref_tangent + var * S.dref;
        double tt = 0;
        for (int b = 0; b < D; ++b) {
          double s = 0;
          for (int a = 0; a < S.dref; ++a) s = fma(tr[a], Jt[a][b], s);
          tt = fma(s, s, tt);
        }
        meas = sqrt(tt);
      } else {
        meas = 1.0;
      }
    }
    *meas_out = meas;
  }
}

template <int D>
__device__ __forceinline__ int sym_index(int b, int d) {
  // components: diagonal first, then (0,1),(0,2),(1,2)
  if (b == d) return b;
  if (D == 2) return 2;
  int lo = b < d ? b : d, hi = b < d ? d : b;
  return lo == 0 ? (hi == 1 ? 3 : 4) : 5;
}

template <int D>
__device__ void eval_op(const DomainDev& dom, const OpInst& O, long long cell, int q, int i,
                        const double* sm_geo, double* out /* stride ndp */) {
  constexpr int S_ = D * (D + 1) / 2;
  const SlotDev& S = dom.slot[O.slot];
  const int nq = dom.nq;
  const int nd = O.ndp;        // component stride of the padded shared-memory rows
  const int var = S.variant ? S.variant[cell] : 0;
  const int qg = dom.nqg == 1 ? 0 : q;
  const double* rec = sm_geo + ((size_t)O.slot * dom.nqg + qg) * GEO_STRIDE;
  const size_t tq = ((size_t)var * nq + q) * S.ndofs + i;
  if (O.op == VLFS_OP_VAL) {
    out[0] = S.val[tq];
    return;
  }
  double g[3] = {0, 0, 0};
  if (O.op == VLFS_OP_GRAD || O.op == VLFS_OP_DDIR || O.op == VLFS_OP_GRAD_NOWN) {
    const double* gr = S.grad + tq * S.dref;
#pragma unroll
    for (int b = 0; b < D; ++b) {
      double s = 0;
      for (int a = 0; a < S.dref; ++a) s = fma(rec[b * 3 + a], gr[a], s);
      g[b] = s;
    }
    if (O.op == VLFS_OP_GRAD) {
#pragma unroll
      for (int b = 0; b < D; ++b) out[(size_t)b * nd] = g[b];
    } else if (O.op == VLFS_OP_DDIR) {
      double s = 0;
#pragma unroll
      for (int b = 0; b < D; ++b) s = fma(g[b], O.dir[b], s);
      out[0] = s;
    } else {
      double s = 0;
#pragma unroll
      for (int b = 0; b < D; ++b) s = fma(g[b], rec[9 + b], s);
      out[0] = s;
    }
    return;
  }
  // second-derivative operators
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  {
    const double* hr = S.hess + tq * S.dref * S.dref;
    double T[3][3];   // T[b][c] = sum_a pinv[b][a] Hr[a][c]
    for (int b = 0; b < D; ++b)
      for (int c = 0; c < S.dref; ++c) {
        double s = 0;
        for (int a = 0; a < S.dref; ++a) s = fma(rec[b * 3 + a], hr[a * S.dref + c], s);
        T[b][c] = s;
      }
    for (int b = 0; b < D; ++b)
      for (int d = 0; d < D; ++d) {
        double s = 0;
        for (int c = 0; c < S.dref; ++c) s = fma(T[b][c], rec[d * 3 + c], s);
        H[b][d] = s;
      }
  }
  if (O.op == VLFS_OP_LAPL) {
    double s = 0;
#pragma unroll
    for (int b = 0; b < D; ++b) s += H[b][b];
    out[0] = s;
    return;
  }
  if (O.op == VLFS_OP_HESS) {
#pragma unroll
    for (int b = 0; b < D; ++b)
#pragma unroll
      for (int d = b; d < D; ++d) out[(size_t)sym_index<D>(b, d) * nd] = H[b][d];
    return;
  }
  // C : H
  double CH[3][3];
  for (int b = 0; b < D; ++b)
    for (int d = 0; d < D; ++d) {
      double s = 0;
      const double* Cbd = dom.C + ((size_t)(b * D + d) * D) * D;
      for (int e = 0; e < D; ++e)
        for (int f = 0; f < D; ++f) s = fma(Cbd[e * D + f], H[e][f], s);
      CH[b][d] = s;
    }
  if (O.op == VLFS_OP_CHESS) {
#pragma unroll
    for (int b = 0; b < D; ++b)
#pragma unroll
      for (int d = b; d < D; ++d)
        out[(size_t)sym_index<D>(b, d) * nd] = (b == d ? 1.0 : 2.0) * CH[b][d];
    (void)S_;
    return;
  }
  // VLFS_OP_CHESS_NPLUS: contraction with the normal of slot 0
  const double* rec0 = sm_geo + (size_t)qg * GEO_STRIDE;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double s = 0;
#pragma unroll
    for (int b = 0; b < D; ++b) s = fma(CH[a][b], rec0[9 + b], s);
    out[(size_t)a * nd] = s;
  }
}

// global-space reduction (no generic-address dispatch, no return value)
__device__ __forceinline__ void red_add(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// add `v` into CSR value `base[stride*slot]`.  Exact zeros (basis functions that vanish on the
// integration face: the explicitly stored zeros of the pattern) are not sent: x + 0 == x, and
// every scattered reduction costs the LSU a full lane slot.
__device__ __forceinline__ void scatter_value(double* base, int stride, int d, double v) {
  if (d >= 0 && v != 0.0) red_add(base + (size_t)stride * d, v);
}

struct AsmPtrs {
  double* mat[4];
  double* vec[4];
};

// Pipelined over the cells of a CTA: iteration k computes the geometry of cell k+2, the
// per-cell operator tables and coefficient fields of cell k+1 and the tiles of cell k, with
// one barrier per iteration, so the dependent global loads of the preparation stages overlap
// the FP64 tile contraction of other warps.
template <int D>
__global__ void __launch_bounds__(256, 3)
integrate_cells(DomainDev dom, const RhsDev* __restrict__ rhs, int nrhs, const int* __restrict__ dest,
                const int* __restrict__ cell_list, long long ncells_run,
                AsmPtrs ptrs, int is_complex, int do_mat, int do_vec, int dest_by_run, unsigned stage_mask,
                int use_tref) {
  extern __shared__ __align__(16) double sm[];
  const int nq = dom.nq, nqg = dom.nqg;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int vstride = is_complex ? 2 : 1;
  const long long first = blockIdx.x, step = gridDim.x;
  auto cell_at = [&](long long k) -> long long {
    const long long r = first + k * step;
    if (r >= ncells_run) return -1;
    return cell_list ? (long long)cell_list[r] : r;
  };
  // 4 geometry buffers: a warp one phase ahead writes cell k+3 while a slow warp still reads cell k
  auto geo_buf = [&](long long k) { return sm + (size_t)(k & 3) * dom.smem_geo_stride; };
  auto stage_geometry = [&](long long cell, double* geo) {
    if (cell < 0) return;
    // (reversed thread order: the serial geometry chain and the coefficient-field loads go to the
    // last warp, whose lanes beyond the tile count have no contraction work)
    for (int t = nth - 1 - tid; t < dom.nslots * nqg; t += nth) {
      const int s = t / nqg, q = t - s * nqg;
      double m;
      slot_geometry<D>(dom.slot[s], cell, q, nq, geo + (size_t)t * GEO_STRIDE,
                       s == dom.measure_slot ? &m : nullptr, dom.measure_kind);
      if (s == dom.measure_slot) {
        if (nqg == 1) {
          for (int qq = 0; qq < nq; ++qq) geo[dom.smem_meas_off + qq] = m * dom.qweights[qq];
          geo[dom.smem_meas_off + nq] = m;       // constant cell measure (reference-matrix terms)
        } else {
          geo[dom.smem_meas_off + q] = m * dom.qweights[q];
        }
      }
    }
  };
  auto stage_tables = [&](long long cell, const double* geo, int par, bool invariant) {
    if (cell < 0) return;
    if (!invariant) {
      double* w = sm + dom.smem_buf_off + (size_t)par * dom.smem_buf_stride;
      for (int t = nth - 1 - tid; t < dom.nweights * nq; t += nth) {
        const int wi = t / nq, q = t - wi * nq;
        w[t] = dom.weights[((size_t)wi * dom.ncells + cell) * nq + q];
      }
    }
    for (int o = 0; o < dom.nopi; ++o) {
      const OpInst& O = dom.opi[o];
      if ((O.buf_stride == 0) != invariant) continue;
      if (!(stage_mask >> o & 1u)) continue;          // vector-only pass: tables of the linear terms
      const int nd = dom.slot[O.slot].ndofs;
      double* base = sm + O.smem_off + (size_t)par * O.buf_stride;
      for (int t = tid; t < nq * nd; t += nth) {
        const int q = t / nd, i = t - q * nd;
        eval_op<D>(dom, O, cell, q, i, geo, base + (size_t)q * O.ncomp * O.ndp + i);
      }
    }
  };

  // operator tables are zero padded to a multiple of 4 local DOFs; the padding is never
  // written again, so clearing once per CTA is enough
  for (int t = dom.smem_inv_off + tid; t < dom.smem_total; t += nth) sm[t] = 0.0;
  stage_geometry(cell_at(0), geo_buf(0));
  __syncthreads();
  // split-phase barrier: a thread arrives when its share of the preparation of cell k+1 is
  // written and waits only before it needs those tables, i.e. after its own tiles of cell k;
  // with triple-buffered tables the tile phases of different warps may skew by a whole phase
  __shared__ cuda::barrier<cuda::thread_scope_block> pipe_bar;
  if (tid == 0) init(&pipe_bar, nth);
  __syncthreads();
  stage_tables(cell_at(0), geo_buf(0), 0, true);     // cell-invariant tables, once
  stage_tables(cell_at(0), geo_buf(0), 0, false);
  stage_geometry(cell_at(1), geo_buf(1));
  __syncthreads();
  // reference matrices of the cell-invariant contractions, once per CTA (not worth it when the
  // CTA integrates a single cell, e.g. class representatives: the untiled precompute costs more
  // than the tiled contraction it saves - ncu source page: 50-60 % of such launches)
  if (do_mat && use_tref && dom.smem_tref_len > 0) {
    const int nprod = dom.target_first[dom.nblocks] > 0
        ? dom.targets[dom.target_first[dom.nblocks] - 1].first + dom.targets[dom.target_first[dom.nblocks] - 1].count : 0;
    for (int pi = 0; pi < nprod; ++pi) {
      const TileProd P = dom.prods[pi];
      if (P.tref < 0 || P.tref_owner != pi) continue;
      const int nr4 = P.nr4, nc4 = P.ldt;
      for (int t = tid; t < nr4 * nc4; t += nth) {
        const int i = t / nc4, j = t - i * nc4;
        double acc = 0.0;
        for (int q = 0; q < nq; ++q) {
          double d = 0.0;
          for (int c = 0; c < P.ncomp; ++c)
            d = fma(sm[P.offA + (q * P.ncomp + c) * P.ndpA + i], sm[P.offB + (q * P.ncomp + c) * P.ndpB + j], d);
          acc = fma(dom.qweights[q], d, acc);
        }
        sm[P.tref + t] = acc;
      }
    }
    __syncthreads();
  }

  for (long long k = 0;; ++k) {
    const long long cell = cell_at(k);
    if (cell < 0) break;
    const int par = (int)(k % dom.nbuf);
    stage_geometry(cell_at(k + 2), geo_buf(k + 2));
    stage_tables(cell_at(k + 1), geo_buf(k + 1), (int)((k + 1) % dom.nbuf), false);
    cuda::barrier<cuda::thread_scope_block>::arrival_token token;
    const bool split = dom.nbuf >= 3;
    if (split) token = pipe_bar.arrive();
    const double* sm_meas = geo_buf(k) + dom.smem_meas_off;
    const double* sm_w = sm + dom.smem_buf_off + (size_t)par * dom.smem_buf_stride;
    // ---- bilinear terms, 4x4 register tiles --------------------------------------
    // Every (block, matrix, re|im) target is a small GEMM  T = A^T diag(w) B over the
    // (point, component) index; a thread owns a 4x4 tile of one block, reads 4+4 operands per
    // contraction index with two 128-bit shared loads each and issues 16 DFMAs.
    if (do_mat) {
      const int ntiles = dom.tile_base[dom.nblocks];
      for (int tile = tid; tile < ntiles; tile += nth) {
        int b = 0;
        while (tile >= dom.tile_base[b + 1]) ++b;
        const int lt = tile - dom.tile_base[b];
        const int ti = lt / dom.tiles_j[b], tj = lt - ti * dom.tiles_j[b];
        const int i0 = ti * 4, j0 = tj * 4;
        // the tile's 16 CSR slots (tile-major dest, -1 = padding), fetched before the math
        const int4* dt = reinterpret_cast<const int4*>(dest) +
                         ((size_t)(dest_by_run ? first + k * step : cell) * ntiles + tile) * 4;
        int slot[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int4 v = __ldg(dt + a);
          slot[a][0] = v.x; slot[a][1] = v.y; slot[a][2] = v.z; slot[a][3] = v.w;
        }
        for (int tg = dom.target_first[b]; tg < dom.target_first[b + 1]; ++tg) {
          const TileTarget T = dom.targets[tg];
          double acc[4][4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
          for (int pi = T.first; pi < T.first + T.count; ++pi) {
            const TileProd P = dom.prods[pi];
            if (use_tref && P.tref >= 0) {  // cell measure x coefficient x reference matrix
              const double w = sm_meas[nq] * P.coef;
              const double* Tr = sm + P.tref + i0 * P.ldt + j0;
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const double2 t01 = *reinterpret_cast<const double2*>(Tr + a * P.ldt);
                const double2 t23 = *reinterpret_cast<const double2*>(Tr + a * P.ldt + 2);
                acc[a][0] = fma(w, t01.x, acc[a][0]); acc[a][1] = fma(w, t01.y, acc[a][1]);
                acc[a][2] = fma(w, t23.x, acc[a][2]); acc[a][3] = fma(w, t23.y, acc[a][3]);
              }
              continue;
            }
            const double* Ap = sm + P.offA + par * P.bufA + i0;
            const double* Bp = sm + P.offB + par * P.bufB + j0;
            const double* W = P.weight >= 0 ? sm_w + P.weight * nq : nullptr;
            const int sa = P.ndpA, sb = P.ndpB, nc = P.ncomp;
            for (int q = 0; q < nq; ++q) {
              double w = sm_meas[q] * P.coef;
              if (W) w *= W[q];
              for (int c = 0; c < nc; ++c) {
                const double2 a01 = *reinterpret_cast<const double2*>(Ap);
                const double2 a23 = *reinterpret_cast<const double2*>(Ap + 2);
                const double2 b01 = *reinterpret_cast<const double2*>(Bp);
                const double2 b23 = *reinterpret_cast<const double2*>(Bp + 2);
                const double av[4] = {a01.x * w, a01.y * w, a23.x * w, a23.y * w};
                const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                  for (int e = 0; e < 4; ++e) acc[a][e] = fma(av[a], bv[e], acc[a][e]);
                Ap += sa; Bp += sb;
              }
            }
          }
          double* out = ptrs.mat[T.mat] + T.part;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              scatter_value(out, vstride, slot[a][e], acc[a][e]);
        }
      }
    }
    // ---- linear terms ---------------------------------------------------------------
    if (do_vec) {
      for (int t = 0; t < nrhs; ++t) {
        const RhsDev& R = rhs[t];
        for (int s = 0; s < dom.nslots; ++s) {
          if (R.opi[s] < 0) continue;
          const int nd = dom.slot[s].ndofs;
          const OpInst& O = dom.opi[R.opi[s]];
          const double* A = sm + O.smem_off + (size_t)par * O.buf_stride;
          const double* Wr = R.wre >= 0 ? sm_w + R.wre * nq : nullptr;
          const double* Wi = R.wim >= 0 ? sm_w + R.wim * nq : nullptr;
          for (int i = tid; i < nd; i += nth) {
            double ar = 0, ai = 0;
            for (int q = 0; q < nq; ++q) {
              double v = sm_meas[q] * A[(size_t)q * O.ndp + i];
              ar = fma(Wr ? Wr[q] : 1.0, v, ar);
              if (Wi) ai = fma(Wi[q], v, ai);
            }
            ar *= R.sc[s]; ai *= R.sc[s];
            double vr = R.cre * ar - R.cim * ai, vi = R.cre * ai + R.cim * ar;
            const int dof = dom.slot[s].dofs[(size_t)cell * nd + i];
            if (is_complex) {
              red_add(ptrs.vec[R.vec] + 2 * (size_t)dof, vr);
              red_add(ptrs.vec[R.vec] + 2 * (size_t)dof + 1, vi);
            } else {
              red_add(ptrs.vec[R.vec] + (size_t)dof, vr);
            }
          }
        }
      }
    }
    if (split) pipe_bar.wait(std::move(token));
    else pipe_bar.arrive_and_wait();       // two table buffers only: full barrier per cell
  }
}

// ---- functionals of FE functions:  sum_cells ∫ | sum_s scale_s OP(u_h) - f |^2 dμ  ------------
// (L2 errors and energies of src/Periodic_Beam.jl:119-120,141-148).  One CTA per cell, one thread
// per quadrature point; u_h is evaluated from the DOF vector x through the same operator code
// as the assembly.  f = coefficient field sampled at the points, or absent.
template <int D>
__global__ void __launch_bounds__(128)
functional_cells(DomainDev dom, OpInst O0, OpInst O1, double sc0, double sc1, int wf,
                 const double* __restrict__ x, double* __restrict__ result) {
  extern __shared__ __align__(16) double sm[];        // geometry records [nslots][nq][12]
  __shared__ double red[4];
  const int nq = dom.nq, tid = threadIdx.x;
  double total = 0.0;
  for (long long cell = blockIdx.x; cell < dom.ncells; cell += gridDim.x) {
    __syncthreads();
    double meas = 0.0;
    for (int t = tid; t < dom.nslots * nq; t += blockDim.x) {
      const int s = t / nq, q = t - s * nq;
      double m;
      slot_geometry<D>(dom.slot[s], cell, q, nq, sm + (size_t)t * GEO_STRIDE,
                       s == dom.measure_slot ? &m : nullptr, dom.measure_kind);
      if (s == dom.measure_slot) sm[(size_t)dom.nslots * nq * GEO_STRIDE + q] = m * dom.qweights[q];
    }
    __syncthreads();
    for (int q = tid; q < nq; q += blockDim.x) {
      double v[6] = {0, 0, 0, 0, 0, 0};
      for (int k = 0; k < 2; ++k) {
        const OpInst& O = k == 0 ? O0 : O1;
        const double sc = k == 0 ? sc0 : sc1;
        if (sc == 0.0) continue;
        const SlotDev& S = dom.slot[O.slot];
        for (int i = 0; i < S.ndofs; ++i) {
          double tmp[6];
          eval_op<D>(dom, O, cell, q, 0 + i, sm, tmp);      // component stride O.ndp == 1
          const double xi = x[S.dofs[(size_t)cell * S.ndofs + i]] * sc;
          for (int c = 0; c < O.ncomp; ++c) v[c] = fma(tmp[c], xi, v[c]);
        }
      }
      if (wf >= 0) v[0] -= dom.weights[((size_t)wf * dom.ncells + cell) * nq + q];
      double s2 = 0.0;
      for (int c = 0; c < O0.ncomp; ++c) s2 = fma(v[c], v[c], s2);
      meas = sm[(size_t)dom.nslots * nq * GEO_STRIDE + q];
      total = fma(meas, s2, total);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_down_sync(0xffffffffu, total, o);
  if ((tid & 31) == 0) red[tid >> 5] = total;
  __syncthreads();
  if (tid == 0) {
    double t = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
    result[1 + blockIdx.x] = t;          // per-CTA partial; functional_finish adds them in a fixed order
  }
}

// result[0] = sum of the nblk per-CTA partials result[1..nblk], in an order that does not depend on
// the schedule (no atomics: the value is bitwise repeatable)
__global__ void __launch_bounds__(256) functional_finish(double* __restrict__ result, int nblk) {
  __shared__ double sh[256];
  double t = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 256) t += result[1 + b];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) result[0] = sh[0];
}

// ---- fast path: ∫ ∇u·∇v on constant-Jacobian cells --------------------------------
// reference matrices, symmetrised:  Sref[m][i][j], m over sym components of dref x dref
template <int DR>
__global__ void gradgrad_reference(SlotDev S, const double* __restrict__ qw, int nq,
                                   double* __restrict__ Sref) {
  constexpr int NS = DR * (DR + 1) / 2;
  const int nd = S.ndofs;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nd * nd; idx += gridDim.x * blockDim.x) {
    int i = idx / nd, j = idx - i * nd;
    double acc[NS];
#pragma unroll
    for (int m = 0; m < NS; ++m) acc[m] = 0;
    for (int q = 0; q < nq; ++q) {
      const double* gi = S.grad + ((size_t)q * nd + i) * DR;
      const double* gj = S.grad + ((size_t)q * nd + j) * DR;
      double w = qw[q];
      int m = 0;
#pragma unroll
      for (int a = 0; a < DR; ++a) { acc[m] = fma(w * gi[a], gj[a], acc[m]); ++m; }
#pragma unroll
      for (int a = 0; a < DR; ++a)
#pragma unroll
        for (int c = a + 1; c < DR; ++c) {
          acc[m] = fma(w, gi[a] * gj[c] + gi[c] * gj[a], acc[m]);
          ++m;
        }
    }
#pragma unroll
    for (int m = 0; m < NS; ++m) Sref[(size_t)m * nd * nd + idx] = acc[m];
  }
}

// ---- reference-matrix contributions gathered per CSR slot -----------------------------------
// Geometry-only per-cell factors (computed once per pattern): measure of the measure slot and,
// for own-cell volume slots, meas * sum_b pinv[b][a] pinv[b][c] in symmetric component order.
template <int D>
__global__ void cell_geometry(DomainDev dom, int want_gradgrad, double* __restrict__ out) {
  constexpr int NS = D * (D + 1) / 2;
  for (long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x; cell < dom.ncells;
       cell += (long long)gridDim.x * blockDim.x) {
    double rec[GEO_STRIDE], meas = 0.0;
    slot_geometry<D>(dom.slot[dom.measure_slot], cell, 0, dom.nq, rec, &meas, dom.measure_kind);
    double* o = out + (size_t)cell * VLFS_CELLGEO;
    o[0] = meas;
    for (int m = 0; m < NS && m < VLFS_CELLGEO - 1; ++m) o[1 + m] = 0.0;
    if (want_gradgrad) {
      double vol = 0.0;
      slot_geometry<D>(dom.slot[0], cell, 0, dom.nq, rec, &vol, VLFS_MEASURE_CELL);
      int m = 0;
      for (int a = 0; a < D; ++a) {
        double t = 0;
        for (int b = 0; b < D; ++b) t = fma(rec[b * 3 + a], rec[b * 3 + a], t);
        o[1 + m++] = vol * t;
      }
      for (int a = 0; a < D; ++a)
        for (int c = a + 1; c < D; ++c) {
          double t = 0;
          for (int b = 0; b < D; ++b) t = fma(rec[b * 3 + a], rec[b * 3 + c], t);
          o[1 + m++] = vol * t;
        }
    }
  }
}

// Linear forms on constant-Jacobian cells whose operators are cell-invariant value tables:
// vec[dof] += coef * sc * meas(cell) * sum_q qw[q] (W_re + i W_im)(cell,q) N_i(q).  One warp per
// cell, lanes over the local DOFs, tables through the read-only path; no shared memory, no barrier
// (the general kernel pays its whole per-cell pipeline for a handful of FMAs here).
__global__ void __launch_bounds__(256)
rhs_affine(DomainDev dom, const RhsDev* __restrict__ rhs, int nrhs, const double* __restrict__ cellgeo,
           AsmPtrs ptrs, int is_complex) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nq = dom.nq;
  for (long long cell = warp; cell < dom.ncells; cell += nwarps) {
    const double meas = cellgeo[(size_t)cell * VLFS_CELLGEO];
    for (int t = 0; t < nrhs; ++t) {
      const RhsDev R = rhs[t];
      const double* Wr = R.wre >= 0 ? dom.weights + ((size_t)R.wre * dom.ncells + cell) * nq : nullptr;
      const double* Wi = R.wim >= 0 ? dom.weights + ((size_t)R.wim * dom.ncells + cell) * nq : nullptr;
      for (int s = 0; s < dom.nslots; ++s) {
        if (R.opi[s] < 0) continue;
        const SlotDev& S = dom.slot[s];
        for (int i = lane; i < S.ndofs; i += 32) {
          double ar = 0.0, ai = 0.0;
          for (int q = 0; q < nq; ++q) {
            const double v = dom.qweights[q] * __ldg(S.val + (size_t)q * S.ndofs + i);
            ar = fma(Wr ? Wr[q] : 1.0, v, ar);
            if (Wi) ai = fma(Wi[q], v, ai);
          }
          ar *= meas * R.sc[s]; ai *= meas * R.sc[s];
          const double vr = R.cre * ar - R.cim * ai, vi = R.cre * ai + R.cim * ar;
          const int dof = S.dofs[(size_t)cell * S.ndofs + i];
          if (is_complex) {
            if (vr != 0.0) red_add(ptrs.vec[R.vec] + 2 * (size_t)dof, vr);
            if (vi != 0.0) red_add(ptrs.vec[R.vec] + 2 * (size_t)dof + 1, vi);
          } else if (vr != 0.0) {
            red_add(ptrs.vec[R.vec] + (size_t)dof, vr);
          }
        }
      }
    }
  }
}

// Jacobian (transposed, [dref][D] padded to 3x3) of every slot at the first point: the complete
// geometric description of a constant-Jacobian cell (classification key, computed once)
template <int D>
__global__ void cell_jacobians(DomainDev dom, double* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < dom.ncells * dom.nslots;
       t += (long long)gridDim.x * blockDim.x) {
    const long long cell = t / dom.nslots;
    const SlotDev& S = dom.slot[(int)(t - cell * dom.nslots)];
    const int var = S.variant ? S.variant[cell] : 0;
    const double* gg = S.geo_grad + ((size_t)var * dom.nq) * S.nv * S.dref;
    const double* X = S.coords + (size_t)cell * S.nv * D;
    double Jt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int v = 0; v < S.nv; ++v)
      for (int a = 0; a < S.dref; ++a)
        for (int b = 0; b < D; ++b) Jt[a * 3 + b] = fma(gg[v * S.dref + a], X[v * D + b], Jt[a * 3 + b]);
    for (int k = 0; k < 9; ++k) out[(size_t)t * 9 + k] = Jt[k];
  }
}

// dest of the class representatives: entry (block, li) of representative r -> ctab_off[b] + r*ne + li
__global__ void class_dest(DomainDev dom, int nreps, GatherDom G, int* __restrict__ tiled) {
  const int ntiles = dom.tile_base[dom.nblocks];
  const long long total = (long long)nreps * ntiles * 16;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / (ntiles * 16);
    const int w = (int)(g - r * (ntiles * 16));
    const int tile = w >> 4, a = (w >> 2) & 3, e = w & 3;
    int b = 0;
    while (tile >= dom.tile_base[b + 1]) ++b;
    const BlockDev& B = dom.blk[b];
    const int lt = tile - dom.tile_base[b];
    const int ti = lt / dom.tiles_j[b], tj = lt - ti * dom.tiles_j[b];
    const int i = ti * 4 + a, j = tj * 4 + e;
    tiled[g] = (i < B.nr && j < B.nc) ? (int)(G.ctab_off[b] + r * (long long)(B.nr * B.nc) + i * B.nc + j)
                                      : VLFS_DEST_PAD;
  }
}

// T[i][j] = sum_q qw[q] valA[q][i] valB[q][j]   (both value tables cell-invariant)
__global__ void mass_reference(const double* __restrict__ valA, int ndA, const double* __restrict__ valB, int ndB,
                               const double* __restrict__ qw, int nq, double* __restrict__ T) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < ndA * ndB; idx += gridDim.x * blockDim.x) {
    const int i = idx / ndB, j = idx - i * ndB;
    double acc = 0.0;
    for (int q = 0; q < nq; ++q) acc = fma(qw[q] * valA[(size_t)q * ndA + i], valB[(size_t)q * ndB + j], acc);
    T[idx] = acc;
  }
}

// sorted COO position -> packed (domain:4 | block:2 | local entry:14 | cell:32) record, once per pattern
__global__ void pack_gather_records(const uint32_t* __restrict__ perm, long long ncoo,
                                    const GatherDom* __restrict__ doms, int ndom,
                                    unsigned long long* __restrict__ rec) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < ncoo;
       p += (long long)gridDim.x * blockDim.x) {
    const long long g = perm[p];
    int d = 0;
    while (d + 1 < ndom && g >= doms[d].coo_hi) ++d;
    const GatherDom& G = doms[d];
    const long long r = g - G.coo_lo;
    const long long e = r / G.epc;
    const int le = (int)(r - e * G.epc);
    int b = 0;
    while (b + 1 < G.nblocks && le >= G.blk[b + 1].off) ++b;
    const unsigned long long li = (unsigned long long)(le - G.blk[b].off);
    rec[p] = ((unsigned long long)d << 60) | ((unsigned long long)b << 58) | (li << 32) | (unsigned long long)e;
  }
}

// ctab[block][class][li][mat][part] = sum_products coef * sum_m cls_geo[class][geo_idx+m] * tab_m[li]
__global__ void build_class_tables(const GatherDom* __restrict__ doms, int d, int nmat, int sc) {
  const GatherDom& G = doms[d];
  for (int b = 0; b < G.nblocks; ++b) {
    const int ne = G.blk[b].nr * G.blk[b].nc;
    const int k0 = G.rp_first[b], k1 = G.rp_first[b + 1];
    if (G.ctab_off[b] < 0) continue;
    const long long total = (long long)G.ncls * ne;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
      const int c = (int)(t / ne), li = (int)(t - (long long)c * ne);
      const double* geo = G.cls_geo + (size_t)c * VLFS_CELLGEO;
      double are[4] = {0, 0, 0, 0}, aim[4] = {0, 0, 0, 0};
      for (int k = k0; k < k1; ++k) {
        const RefProd P = G.rp[k];
        double v = 0.0;
        for (int m = 0; m < P.ntab; ++m) v = fma(geo[P.geo_idx + m], G.tabs[P.tab_off + (size_t)m * ne + li], v);
        are[P.mat] = fma(P.cre, v, are[P.mat]);
        aim[P.mat] = fma(P.cim, v, aim[P.mat]);
      }
      for (int m = 0; m < nmat; ++m) {
        double* out = G.ctab_m[m] + (size_t)(G.ctab_off[b] + t) * sc;
        out[0] = are[m];
        if (sc == 2) out[1] = aim[m];
      }
    }
  }
}

// 64-bit (domain, block, entry, cell) records -> 32-bit index of the entry's value in the
// contiguous class tables (valid when every domain is class-tabulated; 0xffffffff = block without a
// table): the gather then needs one load per COO entry besides the record itself
__global__ void compact_gather_records(const unsigned long long* __restrict__ rec, long long ncoo,
                                       const GatherDom* __restrict__ doms, uint32_t* __restrict__ out) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < ncoo;
       p += (long long)gridDim.x * blockDim.x) {
    const unsigned long long r = rec[p];
    const unsigned d = (unsigned)(r >> 60), b = (unsigned)((r >> 58) & 3u), li = (unsigned)((r >> 32) & 0x3fffu);
    const unsigned e = (unsigned)(r & 0xffffffffull);
    const GatherDom& G = doms[d];
    if (G.ncls > 0 && G.ctab_off[b] >= 0)
      out[p] = (uint32_t)(G.ctab_base + G.ctab_off[b] + (long long)G.cls[e] * (G.blk[b].nr * G.blk[b].nc) + li);
    else
      out[p] = 0xffffffffu;
  }
}

struct GatherSmemDom {
  const double* geo;
  const double* tabs;      // reference tables (global memory)
  int ncls;
  const int* cls;
  const double* ctab_m[4];
  long long ctab_off[VLFS_MAX_BLOCKS];
  int has_ctab[VLFS_MAX_BLOCKS];      // block tabulated per class (reference products or all terms)
  int rp_first[VLFS_MAX_BLOCKS + 1];
  int ne[VLFS_MAX_BLOCKS];
};
constexpr int GATHER_MAX_DOM = 16, GATHER_MAX_RP = 96;

// One thread per CSR slot: sums the reference-type contributions of the slot's COO entries in COO
// order (deterministic, no atomics) and WRITES the value - this pass also replaces the zero fill.
template <int NMAT, bool COMPACT>
__global__ void __launch_bounds__(256, 8)
gather_reference(const uint32_t* __restrict__ run_start, const unsigned long long* __restrict__ rec,
                 const uint32_t* __restrict__ rec32, AsmPtrs ctab_all, long long nnz,
                 const GatherDom* __restrict__ doms, int ndom, int nrp_total, AsmPtrs ptrs, int nmat,
                 int is_complex) {
  // (staging the reference tables in shared memory was measured on B200: it halves the occupancy
  // and is slower than reading them through L1)
  __shared__ GatherSmemDom sd[GATHER_MAX_DOM];
  __shared__ RefProd srp[GATHER_MAX_RP];
  __shared__ int srp_base[GATHER_MAX_DOM];
  if (threadIdx.x == 0) {
    int base = 0;
    for (int d = 0; d < ndom; ++d) {
      sd[d].geo = doms[d].geo; sd[d].tabs = doms[d].tabs;
      sd[d].ncls = doms[d].ncls; sd[d].cls = doms[d].cls;
      for (int m = 0; m < 4; ++m) sd[d].ctab_m[m] = doms[d].ctab_m[m];
      for (int b = 0; b < VLFS_MAX_BLOCKS; ++b) {
        sd[d].ctab_off[b] = doms[d].ctab_off[b];
        sd[d].has_ctab[b] = doms[d].ncls > 0 && doms[d].ctab_off[b] >= 0;
      }
      for (int b = 0; b <= VLFS_MAX_BLOCKS; ++b) sd[d].rp_first[b] = doms[d].rp_first[b] + base;
      for (int b = 0; b < VLFS_MAX_BLOCKS; ++b) sd[d].ne[b] = doms[d].blk[b].nr * doms[d].blk[b].nc;
      srp_base[d] = base;
      const int n = doms[d].rp_first[doms[d].nblocks];
      for (int k = 0; k < n; ++k) srp[base + k] = doms[d].rp[k];
      base += n;
    }
  }
  __syncthreads();
  for (long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x; slot < nnz;
       slot += (long long)gridDim.x * blockDim.x) {
    double are[NMAT], aim[NMAT];
#pragma unroll
    for (int m = 0; m < NMAT; ++m) are[m] = aim[m] = 0.0;
    const uint32_t p1 = run_start[slot + 1];
    // (a 4-way unrolled version that fetches records and class ids ahead was measured slower)
    for (uint32_t p = run_start[slot]; p < p1; ++p) {
      if (COMPACT) {                       // every domain is class-tabulated: record = table index
        const uint32_t r = rec32[p];
        if (r == 0xffffffffu) continue;
        if (is_complex) {
#pragma unroll
          for (int m = 0; m < NMAT; ++m) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(ctab_all.mat[m]) + r);
            are[m] += v.x; aim[m] += v.y;
          }
        } else {
#pragma unroll
          for (int m = 0; m < NMAT; ++m) are[m] += __ldg(ctab_all.mat[m] + r);
        }
        continue;
      }
      const unsigned long long r = rec[p];
      const int d = (int)(r >> 60), b = (int)((r >> 58) & 3u), li = (int)((r >> 32) & 0x3fffu);
      const unsigned e = (unsigned)(r & 0xffffffffull);
      const GatherSmemDom& G = sd[d];
      const int k0 = G.rp_first[b], k1 = G.rp_first[b + 1];
      const int ne = G.ne[b];
      if (G.has_ctab[b]) {                                 // tabulated element block of the class
        const int c = __ldg(G.cls + e);
        const size_t t = (size_t)G.ctab_off[b] + (size_t)c * ne + li;
        if (is_complex) {
#pragma unroll
          for (int m = 0; m < NMAT; ++m) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(G.ctab_m[m]) + t);
            are[m] += v.x; aim[m] += v.y;
          }
        } else {
#pragma unroll
          for (int m = 0; m < NMAT; ++m) are[m] += __ldg(G.ctab_m[m] + t);
        }
        continue;
      }
      if (k0 == k1) continue;
      const double* geo = G.geo + (size_t)e * VLFS_CELLGEO;
      for (int k = k0; k < k1; ++k) {
        const RefProd& P = srp[k];
        const double* tab = G.tabs + P.tab_off + li;
        double v = 0.0;
        for (int m = 0; m < P.ntab; ++m) v = fma(__ldg(geo + P.geo_idx + m), tab[(size_t)m * ne], v);
#pragma unroll
        for (int m = 0; m < NMAT; ++m)          // static indexing keeps the accumulators in registers
          if (P.mat == m) { are[m] = fma(P.cre, v, are[m]); aim[m] = fma(P.cim, v, aim[m]); }
      }
    }
#pragma unroll
    for (int m = 0; m < NMAT; ++m) {
      if (is_complex) reinterpret_cast<double2*>(ptrs.mat[m])[slot] = make_double2(are[m], aim[m]);
      else ptrs.mat[m][slot] = are[m];
    }
  }
}

// Compact gather (every domain class-tabulated: a record is the index of a class-table entry)
// with K slots per thread: a CTA of 128 threads covers 128*K consecutive slots (thread t owns
// slots t, t+128, ...: loads and stores stay coalesced), so K independent record -> table-entry
// load chains are in flight per thread.  Same additions in the same (COO) order as
// gather_reference.  Measured on Y1: 0.62 ms (K = 1) -> 0.50 ms (K = 2).
template <int NMAT, bool CPLX, int K>
__global__ void __launch_bounds__(128)
gather_compact(const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ rec32, AsmPtrs ctab_all,
               long long nnz, AsmPtrs ptrs) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  const long long ntiles = (nnz + 128 * K - 1) / (128 * K);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long s0 = tile * (128 * K) + threadIdx.x;
    T acc[K][NMAT];
    uint32_t p[K], e[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int m = 0; m < NMAT; ++m) acc[k][m] = T{};
      const long long sl = s0 + 128 * k;
      p[k] = e[k] = 0;
      if (sl < nnz) { p[k] = __ldg(run_start + sl); e[k] = __ldg(run_start + sl + 1); }
    }
    for (;;) {
      bool any = false;
      uint32_t r[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const bool live = p[k] < e[k];
        any |= live;
        r[k] = live ? __ldg(rec32 + p[k]) : 0xffffffffu;
        p[k] += live;
      }
      if (!any) break;
      T v[K][NMAT];
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int m = 0; m < NMAT; ++m)
          v[k][m] = r[k] != 0xffffffffu ? __ldg(reinterpret_cast<const T*>(ctab_all.mat[m]) + r[k]) : T{};
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int m = 0; m < NMAT; ++m)
          if (r[k] != 0xffffffffu) {           // (an absent record adds nothing, not even +0.0)
            if constexpr (CPLX) { acc[k][m].x += v[k][m].x; acc[k][m].y += v[k][m].y; }
            else acc[k][m] += v[k][m];
          }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const long long sl = s0 + 128 * k;
      if (sl < nnz) {
#pragma unroll
        for (int m = 0; m < NMAT; ++m) reinterpret_cast<T*>(ptrs.mat[m])[sl] = acc[k][m];
      }
    }
  }
}

}  // namespace

void launch_cell_geometry(const DomainDev& dom, int want_gradgrad, double* out, cudaStream_t s) {
  if (dom.ncells == 0) return;
  unsigned grid = (unsigned)((dom.ncells + 127) / 128);
  if (grid > 148 * 16) grid = 148 * 16;
  if (dom.D == 1) cell_geometry<1><<<grid, 128, 0, s>>>(dom, want_gradgrad, out);
  else if (dom.D == 2) cell_geometry<2><<<grid, 128, 0, s>>>(dom, want_gradgrad, out);
  else cell_geometry<3><<<grid, 128, 0, s>>>(dom, want_gradgrad, out);
  CUDA_TRY(cudaGetLastError());
}

void launch_rhs_affine(const DomainDev& dom, const RhsDev* rhs_dev, int nrhs, const double* cellgeo,
                       double* const* vecs, int nvec, int is_complex, int nsm, cudaStream_t s) {
  if (dom.ncells == 0 || nrhs == 0) return;
  AsmPtrs P{};
  for (int v = 0; v < nvec && v < 4; ++v) P.vec[v] = vecs[v];
  long long grid = (dom.ncells + 7) / 8;
  if (grid > (long long)nsm * 16) grid = (long long)nsm * 16;
  rhs_affine<<<(unsigned)grid, 256, 0, s>>>(dom, rhs_dev, nrhs, cellgeo, P, is_complex);
  CUDA_TRY(cudaGetLastError());
}

void launch_cell_jacobians(const DomainDev& dom, double* out, cudaStream_t s) {
  if (dom.ncells == 0) return;
  unsigned grid = (unsigned)((dom.ncells * dom.nslots + 127) / 128);
  if (grid > 148 * 16) grid = 148 * 16;
  if (dom.D == 1) cell_jacobians<1><<<grid, 128, 0, s>>>(dom, out);
  else if (dom.D == 2) cell_jacobians<2><<<grid, 128, 0, s>>>(dom, out);
  else cell_jacobians<3><<<grid, 128, 0, s>>>(dom, out);
  CUDA_TRY(cudaGetLastError());
}

void launch_class_dest(const DomainDev& dom, int nreps, const GatherDom& G, int* tiled, cudaStream_t s) {
  const long long total = (long long)nreps * dom.tile_base[dom.nblocks] * 16;
  if (total == 0) return;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  class_dest<<<grid, 256, 0, s>>>(dom, nreps, G, tiled);
  CUDA_TRY(cudaGetLastError());
}

void launch_mass_reference(const double* valA, int ndA, const double* valB, int ndB, const double* qw, int nq,
                           double* T, cudaStream_t s) {
  mass_reference<<<(ndA * ndB + 255) / 256, 256, 0, s>>>(valA, ndA, valB, ndB, qw, nq, T);
  CUDA_TRY(cudaGetLastError());
}

void launch_build_class_tables(const GatherDom* doms, int d, long long work, int nmat, int is_complex, cudaStream_t s) {
  if (work == 0) return;
  long long grid = (work + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  build_class_tables<<<(unsigned)grid, 256, 0, s>>>(doms, d, nmat, is_complex ? 2 : 1);
  CUDA_TRY(cudaGetLastError());
}

void launch_pack_gather_records(const uint32_t* perm, long long ncoo, const GatherDom* doms, int ndom,
                                unsigned long long* rec, cudaStream_t s) {
  if (ncoo == 0) return;
  long long grid = (ncoo + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  pack_gather_records<<<(unsigned)grid, 256, 0, s>>>(perm, ncoo, doms, ndom, rec);
  CUDA_TRY(cudaGetLastError());
}

void launch_compact_gather_records(const unsigned long long* rec, long long ncoo, const GatherDom* doms,
                                   uint32_t* out, cudaStream_t s) {
  if (ncoo == 0) return;
  long long grid = (ncoo + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  compact_gather_records<<<(unsigned)grid, 256, 0, s>>>(rec, ncoo, doms, out);
  CUDA_TRY(cudaGetLastError());
}

void launch_gather_reference(const uint32_t* run_start, const unsigned long long* rec, const uint32_t* rec32,
                             double* const* ctab_all, long long nnz, const GatherDom* doms, int ndom,
                             int nrp_total, long long tab_total, double* const* mats, int nmat, int is_complex,
                             int nsm, cudaStream_t s) {
  if (nnz == 0) return;
  if (ndom > GATHER_MAX_DOM || nrp_total > GATHER_MAX_RP)
    throw std::string("gather_reference: too many domains / reference products");
  AsmPtrs P{}, CT{};
  for (int m = 0; m < nmat && m < 4; ++m) { P.mat[m] = mats[m]; CT.mat[m] = ctab_all[m]; }
  (void)tab_total;
  long long grid = (nnz + 255) / 256;
  if (grid > (long long)nsm * 32) grid = (long long)nsm * 32;
  auto launch = [&](auto kern) {
    kern<<<(unsigned)grid, 256, 0, s>>>(run_start, rec, rec32, CT, nnz, doms, ndom, nrp_total, P, nmat, is_complex);
  };
  static const int kslots = getenv("VLFS_GATHER_K") ? atoi(getenv("VLFS_GATHER_K")) : 2;
  if (rec32 && kslots > 1) {
    const int K = kslots >= 4 ? 4 : kslots;
    long long g = (nnz + 128 * K - 1) / (128 * K);
    if (g > (long long)nsm * 64) g = (long long)nsm * 64;
    auto go = [&](auto kern) { kern<<<(unsigned)g, 128, 0, s>>>(run_start, rec32, CT, nnz, P); };
#define VLFS_GC(NM, CX) \
    (K == 2 ? go(gather_compact<NM, CX, 2>) : K == 3 ? go(gather_compact<NM, CX, 3>) : go(gather_compact<NM, CX, 4>))
    switch (nmat * 2 + (is_complex ? 1 : 0)) {
      case 2: VLFS_GC(1, false); break;
      case 3: VLFS_GC(1, true); break;
      case 4: VLFS_GC(2, false); break;
      case 5: VLFS_GC(2, true); break;
      case 6: VLFS_GC(3, false); break;
      case 7: VLFS_GC(3, true); break;
      case 8: VLFS_GC(4, false); break;
      default: VLFS_GC(4, true); break;
    }
#undef VLFS_GC
    CUDA_TRY(cudaGetLastError());
    return;
  }
  if (rec32) {
    switch (nmat) {
      case 1: launch(gather_reference<1, true>); break;
      case 2: launch(gather_reference<2, true>); break;
      case 3: launch(gather_reference<3, true>); break;
      default: launch(gather_reference<4, true>); break;
    }
  } else {
    switch (nmat) {
      case 1: launch(gather_reference<1, false>); break;
      case 2: launch(gather_reference<2, false>); break;
      case 3: launch(gather_reference<3, false>); break;
      default: launch(gather_reference<4, false>); break;
    }
  }
  CUDA_TRY(cudaGetLastError());
}

// ---- host-side launchers ------------------------------------------------------------
void launch_integrate(const DomainDev& dom, const RhsDev* rhs_dev, int nrhs, const int* dest,
                      const int* cell_list, long long ncells_run, double* const* mats,
                      int nmat, double* const* vecs, int nvec, int is_complex, int do_mat,
                      int do_vec, int nsm, cudaStream_t s, int dest_by_run, unsigned stage_mask) {
  if (ncells_run == 0) return;
  AsmPtrs P{};
  for (int m = 0; m < nmat && m < 4; ++m) P.mat[m] = mats[m];
  for (int v = 0; v < nvec && v < 4; ++v) P.vec[v] = vecs[v];
  size_t smem = (size_t)dom.smem_total * sizeof(double);
  // one thread per 4x4 tile of the cell's touched blocks, whole warps, at most 256
  int ntiles = do_mat ? dom.tile_base[dom.nblocks] : 0;
  // ... or one thread per (point, local DOF) entry of the largest per-cell operator table, whichever is
  // more: with few tiles and second-derivative tables (MultiGeo plate: 16 tiles, 240 Hessian entries of
  // ~100 flops each) the staging is the work, and 64 threads left 2 warps per SM (ncu: FP64 pipe 4 %)
  int staging = 0;
  for (int o = 0; o < dom.nopi; ++o)
    if (dom.opi[o].buf_stride != 0 && (stage_mask >> o & 1u) &&
        (dom.opi[o].op == VLFS_OP_HESS || dom.opi[o].op == VLFS_OP_CHESS || dom.opi[o].op == VLFS_OP_CHESS_NPLUS || dom.opi[o].op == VLFS_OP_LAPL))
      staging = std::max(staging, dom.nq * dom.slot[dom.opi[o].slot].ndofs);     // (value / gradient tables are cheap to stage)
  int threads = ((std::max(ntiles, staging) + 31) / 32) * 32;
  if (threads < 64) threads = 64;
  if (threads > 256) threads = 256;
  if (const char* e = getenv("VLFS_ASM_THREADS")) {      // tuning hook
    int t = atoi(e);
    if (t >= 32 && t <= 256 && t % 32 == 0) threads = t;
  }
  long long grid = ncells_run;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  int by_threads = 2048 / threads;
  if (per_sm > by_threads) per_sm = by_threads;
  int by_regs = 65536 / (threads * 80);          // __launch_bounds__(256,3): <= 80 registers
  if (by_regs < 1) by_regs = 1;
  if (per_sm > by_regs) per_sm = by_regs;
  // persistent CTAs: exactly the resident set, each pipelining over its cells
  long long cap = (long long)nsm * per_sm;
  if (grid > cap && !getenv("VLFS_ASM_ONE_CELL_PER_CTA")) grid = cap;       // (diagnostic switch: no pipelining over cells)
  const int use_tref = ncells_run >= 2 * grid;      // at least two cells per CTA
  auto launch = [&](auto kern) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, threads, smem, s>>>(dom, rhs_dev, nrhs, dest, cell_list, ncells_run, P,
                                               is_complex, do_mat, do_vec, dest_by_run, stage_mask, use_tref);
  };
  if (dom.D == 1) launch(integrate_cells<1>);
  else if (dom.D == 2) launch(integrate_cells<2>);
  else launch(integrate_cells<3>);
  CUDA_TRY(cudaGetLastError());
}

void launch_functional(const DomainDev& dom_in, const OpInst& O0, const OpInst& O1, double sc0, double sc1,
                       int wf, const double* x, double* result, int nsm, cudaStream_t s) {
  if (dom_in.ncells == 0) return;
  DomainDev dom = dom_in;
  dom.nqg = dom.nq;                       // per-point geometry records, whatever the domain's flag
  size_t smem = ((size_t)dom.nslots * dom.nq * GEO_STRIDE + dom.nq) * sizeof(double);
  long long grid = dom.ncells < (long long)nsm * 8 ? dom.ncells : (long long)nsm * 8;
  auto launch = [&](auto kern) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, 128, smem, s>>>(dom, O0, O1, sc0, sc1, wf, x, result);
  };
  if (dom.D == 1) launch(functional_cells<1>);
  else if (dom.D == 2) launch(functional_cells<2>);
  else launch(functional_cells<3>);
  functional_finish<<<1, 256, 0, s>>>(result, (int)grid);
  CUDA_TRY(cudaGetLastError());
}

void launch_gradgrad_reference(const DomainDev& dom, double* Sref, cudaStream_t s) {
  const SlotDev& S = dom.slot[0];
  int ne = S.ndofs * S.ndofs;
  unsigned grid = (ne + 255) / 256;
  if (S.dref == 1) gradgrad_reference<1><<<grid, 256, 0, s>>>(S, dom.qweights, dom.nq, Sref);
  else if (S.dref == 2) gradgrad_reference<2><<<grid, 256, 0, s>>>(S, dom.qweights, dom.nq, Sref);
  else gradgrad_reference<3><<<grid, 256, 0, s>>>(S, dom.qweights, dom.nq, Sref);
  CUDA_TRY(cudaGetLastError());
}

