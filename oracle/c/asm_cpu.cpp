// asm_cpu.cpp - compiled CPU restatement of the assembly path (TEST INFRASTRUCTURE / CPU BASELINE).
//
// The same algorithm class as Gridap's `AffineFEOperator(a,l,X,Y)`: integrate every cell's element
// blocks by quadrature from the term descriptors (the semantics of include/vlfs_b200.h, restated
// from oracle/descriptor_eval.py which the -m "not gpu" tests check against the literal weak forms
// of src/*.jl), fill COO triplets in Gridap's order (domain, cell, block, i, j), then build the
// compressed matrix by a counting sort on rows, a sort of every row by column and the summation of
// duplicates - what `sparse(I,J,V)` does.  Host threads split the cells and the rows; results do
// not depend on their number.  Nothing in monolithicfemvlfs.jl_b200/ links or loads this file:
// only tests/ and bench.py's CPU legs do (through oracle/c_assembly.py).
#include "vlfs_b200.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Cplx { double re, im; };

unsigned nthreads_eff(int requested) {
  if (requested >= 1) return (unsigned)requested;
  unsigned t = std::thread::hardware_concurrency();
  return t < 1 ? 1 : t;
}

template <typename F>
void parallel_ranges(long long n, unsigned nt, F&& f) {
  if (nt <= 1 || n < 2) { f(0LL, n, 0u); return; }
  nt = (unsigned)std::min<long long>(nt, n);
  std::vector<std::thread> th;
  const long long chunk = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; ++t) {
    const long long b = t * chunk, e = std::min(n, b + chunk);
    if (b >= e) break;
    th.emplace_back([&f, b, e, t]() { f(b, e, t); });
  }
  for (auto& x : th) x.join();
}

// inverse and determinant of an n x n matrix, n <= 3 (row-major)
double inv_small(int n, const double* A, double* I) {
  if (n == 1) { I[0] = 1.0 / A[0]; return A[0]; }
  if (n == 2) {
    const double d = A[0] * A[3] - A[1] * A[2], r = 1.0 / d;
    I[0] = A[3] * r; I[1] = -A[1] * r; I[2] = -A[2] * r; I[3] = A[0] * r;
    return d;
  }
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double d = A[0] * c00 + A[1] * c01 + A[2] * c02, r = 1.0 / d;
  I[0] = c00 * r; I[1] = (A[2] * A[7] - A[1] * A[8]) * r; I[2] = (A[1] * A[5] - A[2] * A[4]) * r;
  I[3] = c01 * r; I[4] = (A[0] * A[8] - A[2] * A[6]) * r; I[5] = (A[2] * A[3] - A[0] * A[5]) * r;
  I[6] = c02 * r; I[7] = (A[1] * A[6] - A[0] * A[7]) * r; I[8] = (A[0] * A[4] - A[1] * A[3]) * r;
  return d;
}

struct SlotGeo {               // per quadrature point
  double pinv[9];              // [b][a]  D x dref
  double normal[3];
  double meas, fmeas;
};

void slot_geometry(const vlfs_domain_desc& d, const vlfs_slot_desc& s, long long e, int q, SlotGeo& g) {
  const int D = d.D, dref = s.dref, nv = s.nv;
  const int var = s.variant ? s.variant[e] : 0;
  const double* gg = s.geo_grad + ((size_t)var * d.nq + q) * nv * dref;      // [v][a]
  const double* X = s.coords + (size_t)e * nv * D;                          // [v][b]
  double Jt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};                               // [a][b] dref x D
  for (int v = 0; v < nv; ++v)
    for (int a = 0; a < dref; ++a)
      for (int b = 0; b < D; ++b) Jt[a * D + b] += gg[v * dref + a] * X[v * D + b];
  g.fmeas = 1.0;
  g.normal[0] = g.normal[1] = g.normal[2] = 0.0;
  if (dref == 0) {
    g.meas = 1.0;
  } else if (dref == D) {
    double I[9];
    g.meas = std::fabs(inv_small(D, Jt, I));
    for (int b = 0; b < D; ++b)
      for (int a = 0; a < dref; ++a) g.pinv[b * dref + a] = I[b * D + a];
  } else {
    double G[9], Gi[9];
    for (int a = 0; a < dref; ++a)
      for (int c = 0; c < dref; ++c) {
        double t = 0;
        for (int b = 0; b < D; ++b) t += Jt[a * D + b] * Jt[c * D + b];
        G[a * dref + c] = t;
      }
    g.meas = std::sqrt(inv_small(dref, G, Gi));
    for (int b = 0; b < D; ++b)
      for (int a = 0; a < dref; ++a) {
        double t = 0;
        for (int c = 0; c < dref; ++c) t += Jt[c * D + b] * Gi[c * dref + a];
        g.pinv[b * dref + a] = t;
      }
  }
  if (s.ref_normal) {
    const double* rn = s.ref_normal + (size_t)var * dref;
    double nn = 0;
    for (int b = 0; b < D; ++b) {
      double t = 0;
      for (int a = 0; a < dref; ++a) t += g.pinv[b * dref + a] * rn[a];
      g.normal[b] = t; nn += t * t;
    }
    nn = std::sqrt(nn);
    for (int b = 0; b < D; ++b) g.normal[b] /= nn;
  }
  if (s.ref_tangent) {
    const double* rt = s.ref_tangent + (size_t)var * dref;
    double nn = 0;
    for (int b = 0; b < D; ++b) {
      double t = 0;
      for (int a = 0; a < dref; ++a) t += rt[a] * Jt[a * D + b];
      nn += t * t;
    }
    g.fmeas = std::sqrt(nn);
  }
}

int op_ncomp(int op, int D) {
  switch (op) {
    case VLFS_OP_GRAD: case VLFS_OP_CHESS_NPLUS: return D;
    case VLFS_OP_HESS: case VLFS_OP_CHESS: return D * (D + 1) / 2;
    default: return 1;
  }
}

// rows of one operator of slot k on one cell: out[q][c][i]
void eval_op(const vlfs_domain_desc& d, int k, int op, const double* dir, long long e,
             const std::vector<SlotGeo>& geo /* [slot][q] */, std::vector<double>& out) {
  const vlfs_slot_desc& s = d.slot[k];
  const int D = d.D, dref = s.dref, nd = s.ndofs, nq = d.nq;
  const int nc = op_ncomp(op, D);
  const int var = s.variant ? s.variant[e] : 0;
  out.assign((size_t)nq * nc * nd, 0.0);
  int comps[6][2], ncs = 0;
  for (int a = 0; a < D; ++a) { comps[ncs][0] = a; comps[ncs][1] = a; ++ncs; }
  for (int a = 0; a < D; ++a)
    for (int b = a + 1; b < D; ++b) { comps[ncs][0] = a; comps[ncs][1] = b; ++ncs; }
  for (int q = 0; q < nq; ++q) {
    const SlotGeo& g = geo[(size_t)k * nq + q];
    double* o = out.data() + (size_t)q * nc * nd;
    if (op == VLFS_OP_VAL) {
      const double* v = s.val + ((size_t)var * nq + q) * nd;
      for (int i = 0; i < nd; ++i) o[i] = v[i];
      continue;
    }
    const double* gr = s.grad + ((size_t)var * nq + q) * nd * dref;         // [i][a]
    if (op == VLFS_OP_GRAD || op == VLFS_OP_DDIR || op == VLFS_OP_GRAD_NOWN) {
      for (int i = 0; i < nd; ++i) {
        double gphys[3] = {0, 0, 0};
        for (int b = 0; b < D; ++b)
          for (int a = 0; a < dref; ++a) gphys[b] += g.pinv[b * dref + a] * gr[i * dref + a];
        if (op == VLFS_OP_GRAD) for (int b = 0; b < D; ++b) o[b * nd + i] = gphys[b];
        else {
          const double* w = op == VLFS_OP_DDIR ? dir : g.normal;
          double t = 0;
          for (int b = 0; b < D; ++b) t += gphys[b] * w[b];
          o[i] = t;
        }
      }
      continue;
    }
    const double* hs = s.hess + ((size_t)var * nq + q) * nd * dref * dref;  // [i][a][c]
    const double* npl = geo[(size_t)0 * nq + q].normal;                     // n⁺ = normal of slot 0
    for (int i = 0; i < nd; ++i) {
      double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};                            // [b][d]
      for (int b = 0; b < D; ++b)
        for (int dd = 0; dd < D; ++dd) {
          double t = 0;
          for (int a = 0; a < dref; ++a)
            for (int c = 0; c < dref; ++c)
              t += g.pinv[b * dref + a] * hs[(i * dref + a) * dref + c] * g.pinv[dd * dref + c];
          H[b * D + dd] = t;
        }
      if (op == VLFS_OP_LAPL) {
        double t = 0;
        for (int b = 0; b < D; ++b) t += H[b * D + b];
        o[i] = t;
      } else if (op == VLFS_OP_HESS) {
        for (int c = 0; c < ncs; ++c) o[c * nd + i] = H[comps[c][0] * D + comps[c][1]];
      } else {
        double CH[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int a = 0; a < D; ++a)
          for (int b = 0; b < D; ++b) {
            double t = 0;
            for (int c = 0; c < D; ++c)
              for (int dd = 0; dd < D; ++dd) t += d.C[((a * D + b) * D + c) * D + dd] * H[c * D + dd];
            CH[a * D + b] = t;
          }
        if (op == VLFS_OP_CHESS) {
          for (int c = 0; c < ncs; ++c)
            o[c * nd + i] = (comps[c][0] == comps[c][1] ? 1.0 : 2.0) * CH[comps[c][0] * D + comps[c][1]];
        } else {                                                            // VLFS_OP_CHESS_NPLUS
          for (int a = 0; a < D; ++a) {
            double t = 0;
            for (int b = 0; b < D; ++b) t += CH[a * D + b] * npl[b];
            o[a * nd + i] = t;
          }
        }
      }
    }
  }
}

struct OpKey { int slot, op; double dir[3]; };

struct DomainPlan {
  long long coo_base = 0;
  int entries_per_cell = 0;
  int block_off[VLFS_MAX_SLOTS][VLFS_MAX_SLOTS];       // -1 = block not touched
};

}  // namespace

extern "C" {

typedef struct {
  const vlfs_domain_desc* desc;
  int32_t nterms;
  const vlfs_term_desc* terms;
  int32_t nrhs;
  const vlfs_rhs_desc* rhs;
} oracle_domain;

void oracle_free(void* p) { free(p); }

// Assembles `nmat` matrices on one union pattern (CSR, 0-based, columns ascending, explicit zeros of
// touched blocks kept) and `nvec` vectors.  vals[m] / rowptr / colind are malloc'ed here (free with
// oracle_free); vecs is caller-allocated [nvec][ndofs] (x2 when complex) and overwritten.
// ms[0..2] = integration + COO fill, compression (sort / merge), total.  Returns 0 or -1.
int oracle_assemble(int32_t ndom, const oracle_domain* doms, int64_t ndofs, int32_t is_complex, int32_t nmat,
                    int32_t nvec, int32_t nthreads, int64_t* nnz_out, int64_t* ncoo_out, int32_t** rowptr_out,
                    int32_t** colind_out, double** vals_out, double* vecs, double* ms) {
  const unsigned nt = nthreads_eff(nthreads);
  const int sc = is_complex ? 2 : 1;
  auto t0 = std::chrono::steady_clock::now();
  // ---- COO layout -------------------------------------------------------------------------
  std::vector<DomainPlan> plan(ndom);
  long long ncoo = 0;
  for (int di = 0; di < ndom; ++di) {
    const vlfs_domain_desc& d = *doms[di].desc;
    DomainPlan& P = plan[di];
    P.coo_base = ncoo;
    int off = 0;
    for (int a = 0; a < d.nslots; ++a)
      for (int b = 0; b < d.nslots; ++b) {
        bool touched = false;
        for (int t = 0; t < doms[di].nterms; ++t)
          touched |= doms[di].terms[t].test_scale[a] != 0.0 && doms[di].terms[t].trial_scale[b] != 0.0;
        P.block_off[a][b] = touched ? off : -1;
        if (touched) off += d.slot[a].ndofs * d.slot[b].ndofs;
      }
    P.entries_per_cell = off;
    ncoo += d.ncells * (long long)off;
  }
  std::vector<int> crow((size_t)ncoo), ccol((size_t)ncoo);
  std::vector<std::vector<double>> cval(nmat, std::vector<double>((size_t)ncoo * sc));
  for (long long k = 0; k < (long long)nvec * ndofs * sc; ++k) vecs[k] = 0.0;
  // ---- integration ------------------------------------------------------------------------
  for (int di = 0; di < ndom; ++di) {
    const oracle_domain& od = doms[di];
    const vlfs_domain_desc& d = *od.desc;
    const DomainPlan& P = plan[di];
    if (d.ncells == 0) continue;
    int rhs_nd = 0;
    for (int a = 0; a < d.nslots; ++a) rhs_nd += d.slot[a].ndofs;
    std::vector<Cplx> rbuf(od.nrhs ? (size_t)d.ncells * od.nrhs * rhs_nd : 0, Cplx{0, 0});
    parallel_ranges(d.ncells, nt, [&](long long e0, long long e1, unsigned) {
      const int nq = d.nq, D = d.D;
      std::vector<SlotGeo> geo((size_t)d.nslots * nq);
      std::vector<double> dV(nq);
      std::vector<OpKey> keys;
      std::vector<std::vector<double>> tabs;
      tabs.reserve(64);                       // references into `tabs` stay valid while it grows
      std::vector<Cplx> blk;
      auto table = [&](int slot, int op, const double* dir, long long e) -> const std::vector<double>& {
        for (size_t k = 0; k < keys.size(); ++k)
          if (keys[k].slot == slot && keys[k].op == op &&
              (op != VLFS_OP_DDIR || (keys[k].dir[0] == dir[0] && keys[k].dir[1] == dir[1] && keys[k].dir[2] == dir[2])))
            return tabs[k];
        keys.push_back(OpKey{slot, op, {dir[0], dir[1], dir[2]}});
        if (keys.size() > 64) abort();          // (at most 2 slots x 8 operators x a few directions)
        if (tabs.size() < keys.size()) tabs.emplace_back();
        eval_op(d, slot, op, dir, e, geo, tabs[keys.size() - 1]);
        return tabs[keys.size() - 1];
      };
      for (long long e = e0; e < e1; ++e) {
        keys.clear();
        for (int s = 0; s < d.nslots; ++s)
          for (int q = 0; q < nq; ++q) slot_geometry(d, d.slot[s], e, q, geo[(size_t)s * nq + q]);
        for (int q = 0; q < nq; ++q) {
          const SlotGeo& g = geo[(size_t)d.measure_slot * nq + q];
          const double m = d.measure_kind == VLFS_MEASURE_FACET ? (d.slot[d.measure_slot].ref_tangent ? g.fmeas : 1.0)
                                                               : g.meas;
          dV[q] = m * d.qweights[q];
        }
        for (int a = 0; a < d.nslots; ++a)
          for (int b = 0; b < d.nslots; ++b) {
            if (P.block_off[a][b] < 0) continue;
            const int nr = d.slot[a].ndofs, ncb = d.slot[b].ndofs;
            blk.assign((size_t)nmat * nr * ncb, Cplx{0, 0});
            for (int t = 0; t < od.nterms; ++t) {
              const vlfs_term_desc& T = od.terms[t];
              if (T.test_scale[a] == 0.0 || T.trial_scale[b] == 0.0) continue;
              const std::vector<double>& A = table(a, T.test_op, T.test_dir, e);
              const std::vector<double>& B = table(b, T.trial_op, T.trial_dir, e);
              const int nc = op_ncomp(T.test_op, D);
              const double f = T.test_scale[a] * T.trial_scale[b];
              const double cre = T.coef_re * f, cim = T.coef_im * f;
              const double* W = T.weight >= 0 ? d.weights + ((size_t)T.weight * d.ncells + e) * nq : nullptr;
              Cplx* out = blk.data() + (size_t)T.mat * nr * ncb;
              for (int i = 0; i < nr; ++i)
                for (int j = 0; j < ncb; ++j) {
                  double acc = 0.0;
                  for (int q = 0; q < nq; ++q) {
                    double dot = 0.0;
                    for (int c = 0; c < nc; ++c)
                      dot += A[((size_t)q * nc + c) * nr + i] * B[((size_t)q * nc + c) * ncb + j];
                    acc += (W ? dV[q] * W[q] : dV[q]) * dot;
                  }
                  out[(size_t)i * ncb + j].re += cre * acc;
                  out[(size_t)i * ncb + j].im += cim * acc;
                }
            }
            const long long base = P.coo_base + e * (long long)P.entries_per_cell + P.block_off[a][b];
            const int* ra = d.slot[a].dofs + (size_t)e * nr;
            const int* cb = d.slot[b].dofs + (size_t)e * ncb;
            for (int i = 0; i < nr; ++i)
              for (int j = 0; j < ncb; ++j) {
                const long long p = base + (long long)i * ncb + j;
                crow[p] = ra[i]; ccol[p] = cb[j];
                for (int m = 0; m < nmat; ++m) {
                  const Cplx v = blk[((size_t)m * nr + i) * ncb + j];
                  if (is_complex) { cval[m][2 * p] = v.re; cval[m][2 * p + 1] = v.im; }
                  else cval[m][p] = v.re;
                }
              }
          }
        // linear forms of the cell
        for (int r = 0; r < od.nrhs; ++r) {
          const vlfs_rhs_desc& R = od.rhs[r];
          int off = 0;
          for (int a = 0; a < d.nslots; ++a) {
            const int nd = d.slot[a].ndofs;
            if (R.scale[a] != 0.0) {
              const std::vector<double>& A = table(a, R.op, R.dir, e);
              const double* Wr = R.weight_re >= 0 ? d.weights + ((size_t)R.weight_re * d.ncells + e) * nq : nullptr;
              const double* Wi = R.weight_im >= 0 ? d.weights + ((size_t)R.weight_im * d.ncells + e) * nq : nullptr;
              for (int i = 0; i < nd; ++i) {
                double ar = 0, ai = 0;
                for (int q = 0; q < nq; ++q) {
                  const double v = dV[q] * A[(size_t)q * nd + i];
                  ar += (Wr ? Wr[q] : 1.0) * v;
                  if (Wi) ai += Wi[q] * v;
                }
                ar *= R.scale[a]; ai *= R.scale[a];
                Cplx& o = rbuf[((size_t)e * od.nrhs + r) * rhs_nd + off + i];
                o.re = R.coef_re * ar - R.coef_im * ai;
                o.im = R.coef_re * ai + R.coef_im * ar;
              }
            }
            off += nd;
          }
        }
      }
    });
    // ordered accumulation of the linear forms (cells ascending: independent of the threads)
    for (int r = 0; r < od.nrhs; ++r) {
      const vlfs_rhs_desc& R = od.rhs[r];
      double* vec = vecs + (size_t)R.vec * ndofs * sc;
      for (long long e = 0; e < d.ncells; ++e) {
        int off = 0;
        for (int a = 0; a < d.nslots; ++a) {
          const int nd = d.slot[a].ndofs;
          if (R.scale[a] != 0.0)
            for (int i = 0; i < nd; ++i) {
              const Cplx v = rbuf[((size_t)e * od.nrhs + r) * rhs_nd + off + i];
              const int dof = d.slot[a].dofs[(size_t)e * nd + i];
              if (is_complex) { vec[2 * (size_t)dof] += v.re; vec[2 * (size_t)dof + 1] += v.im; }
              else vec[dof] += v.re;
            }
          off += nd;
        }
      }
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  // ---- sparse(I,J,V): counting sort on rows (stable), per-row sort by column, sum duplicates ----
  std::vector<long long> rb(nt + 1);
  for (unsigned t = 0; t <= nt; ++t) rb[t] = ncoo / nt * t;
  rb[nt] = ncoo;
  std::vector<std::vector<int>> hist(nt, std::vector<int>());
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        hist[t].assign(ndofs, 0);
        for (long long p = rb[t]; p < rb[t + 1]; ++p) ++hist[t][crow[p]];
      });
    for (auto& x : th) x.join();
  }
  std::vector<long long> bstart(ndofs + 1, 0);
  {
    long long acc = 0;
    for (long long r = 0; r < ndofs; ++r) {
      bstart[r] = acc;
      for (unsigned t = 0; t < nt; ++t) { const int h = hist[t][r]; hist[t][r] = (int)(acc - bstart[r]); acc += h; }
    }
    bstart[ndofs] = acc;
  }
  std::vector<long long> order((size_t)ncoo);            // COO index of every bucket entry, COO order kept
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        int* fill = hist[t].data();
        for (long long p = rb[t]; p < rb[t + 1]; ++p) { const int r = crow[p]; order[bstart[r] + fill[r]++] = p; }
      });
    for (auto& x : th) x.join();
  }
  std::vector<std::vector<int>>().swap(hist);
  std::vector<int> rowcount(ndofs + 1, 0);
  parallel_ranges(ndofs, nt, [&](long long r0, long long r1, unsigned) {
    for (long long r = r0; r < r1; ++r) {
      long long* b = order.data() + bstart[r];
      long long* e = order.data() + bstart[r + 1];
      std::stable_sort(b, e, [&](long long x, long long y) { return ccol[x] < ccol[y]; });
      int cnt = 0, prev = -1;
      for (long long* p = b; p < e; ++p) { if (ccol[*p] != prev) { ++cnt; prev = ccol[*p]; } }
      rowcount[r + 1] = cnt;
    }
  });
  int32_t* rowptr = (int32_t*)malloc((size_t)(ndofs + 1) * sizeof(int32_t));
  long long nnz = 0;
  rowptr[0] = 0;
  for (long long r = 0; r < ndofs; ++r) { nnz += rowcount[r + 1]; rowptr[r + 1] = (int32_t)nnz; }
  int32_t* colind = (int32_t*)malloc((size_t)std::max<long long>(nnz, 1) * sizeof(int32_t));
  for (int m = 0; m < nmat; ++m) vals_out[m] = (double*)malloc((size_t)std::max<long long>(nnz, 1) * sc * sizeof(double));
  parallel_ranges(ndofs, nt, [&](long long r0, long long r1, unsigned) {
    for (long long r = r0; r < r1; ++r) {
      long long slot = rowptr[r] - 1;
      int prev = -1;
      for (long long q = bstart[r]; q < bstart[r + 1]; ++q) {
        const long long p = order[q];
        if (ccol[p] != prev) {
          ++slot; prev = ccol[p]; colind[slot] = prev;
          for (int m = 0; m < nmat; ++m)
            for (int k = 0; k < sc; ++k) vals_out[m][slot * sc + k] = cval[m][p * sc + k];
        } else {
          for (int m = 0; m < nmat; ++m)
            for (int k = 0; k < sc; ++k) vals_out[m][slot * sc + k] += cval[m][p * sc + k];
        }
      }
    }
  });
  auto t2 = std::chrono::steady_clock::now();
  *nnz_out = nnz; *ncoo_out = ncoo; *rowptr_out = rowptr; *colind_out = colind;
  if (ms) {
    ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    ms[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    ms[2] = std::chrono::duration<double, std::milli>(t2 - t0).count();
  }
  return 0;
}

// y = A x for a CSR matrix (real, or complex interleaved), rows split over the host threads; run
// `reps` times; returns the mean milliseconds per product (the CPU side of the SpMV metric).
double oracle_spmv(int64_t n, const int32_t* rowptr, const int32_t* colind, const double* vals, int32_t is_complex,
                   const double* x, double* y, int32_t reps, int32_t nthreads) {
  const unsigned nt = nthreads_eff(nthreads);
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; ++r)
    parallel_ranges(n, nt, [&](long long r0, long long r1, unsigned) {
      for (long long i = r0; i < r1; ++i) {
        if (is_complex) {
          double sr = 0, si = 0;
          for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
            const double ar = vals[2 * (size_t)p], ai = vals[2 * (size_t)p + 1];
            const double xr = x[2 * (size_t)colind[p]], xi = x[2 * (size_t)colind[p] + 1];
            sr += ar * xr - ai * xi; si += ar * xi + ai * xr;
          }
          y[2 * i] = sr; y[2 * i + 1] = si;
        } else {
          double sacc = 0;
          for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) sacc += vals[p] * x[colind[p]];
          y[i] = sacc;
        }
      }
    });
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / (reps > 0 ? reps : 1);
}

}  // extern "C"
