// cabi_harness.cpp - calls the C ABI of libvlfs_b200 the way Julia's `ccall` does: plain C types,
// caller-owned buffers, status codes, no C++ or torch objects across the boundary.
//
//   cabi_harness --layout      prints sizeof / offsetof of every struct of include/vlfs_b200.h (CPU only;
//                              tests/test_cabi.py compares them with the ctypes mirror and the Julia shim)
//   cabi_harness --run         needs an sm_100 GPU: (1) 1-D P1 stiffness + mass assembled from hand-written
//                              tables against the analytic tridiagonal matrix, SpMV, solve; (2) an imported
//                              CSR matrix solved without coordinates; (3) the same chain split over two ranks
//                              that share device 0 (vlfs_multi_create: threads + in-process transport).
//                              Exit code 0 = all checks passed, 77 = no usable GPU.
#include "../include/vlfs_b200.h"
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <vector>

#define CHECK(call)                                                                        \
  do {                                                                                     \
    int rc_ = (call);                                                                      \
    if (rc_ != VLFS_OK) {                                                                  \
      fprintf(stderr, "%s:%d: %s -> status %d (%s)\n", __FILE__, __LINE__, #call, rc_, ctx ? vlfs_last_error(ctx) : ""); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

static int layout() {
  printf("vlfs_slot_desc %zu coords %zu dofs %zu ref_tangent %zu\n", sizeof(vlfs_slot_desc), offsetof(vlfs_slot_desc, coords),
         offsetof(vlfs_slot_desc, dofs), offsetof(vlfs_slot_desc, ref_tangent));
  printf("vlfs_domain_desc %zu ncells %zu qweights %zu slot %zu\n", sizeof(vlfs_domain_desc), offsetof(vlfs_domain_desc, ncells),
         offsetof(vlfs_domain_desc, qweights), offsetof(vlfs_domain_desc, slot));
  printf("vlfs_term_desc %zu coef_re %zu test_op %zu test_scale %zu trial_dir %zu\n", sizeof(vlfs_term_desc),
         offsetof(vlfs_term_desc, coef_re), offsetof(vlfs_term_desc, test_op), offsetof(vlfs_term_desc, test_scale),
         offsetof(vlfs_term_desc, trial_dir));
  printf("vlfs_rhs_desc %zu coef_re %zu op %zu scale %zu dir %zu\n", sizeof(vlfs_rhs_desc), offsetof(vlfs_rhs_desc, coef_re),
         offsetof(vlfs_rhs_desc, op), offsetof(vlfs_rhs_desc, scale), offsetof(vlfs_rhs_desc, dir));
  printf("vlfs_solver_opts %zu rtol %zu block_size_hint %zu\n", sizeof(vlfs_solver_opts), offsetof(vlfs_solver_opts, rtol),
         offsetof(vlfs_solver_opts, block_size_hint));
  printf("vlfs_solve_stats %zu rel_residual %zu solve_ms %zu\n", sizeof(vlfs_solve_stats), offsetof(vlfs_solve_stats, rel_residual),
         offsetof(vlfs_solve_stats, solve_ms));
  printf("vlfs_functional_desc %zu scale %zu dir %zu\n", sizeof(vlfs_functional_desc), offsetof(vlfs_functional_desc, scale),
         offsetof(vlfs_functional_desc, dir));
  return 0;
}

// tables of a 1-D P1 mesh of n segments on [0,1] with a 2-point Gauss rule; cells [c0, c1)
struct Mesh1D {
  int nq = 2, nv = 2, nd = 2;
  std::vector<double> coords, qw, val, grad, geo_grad;
  std::vector<int32_t> dofs;
  Mesh1D(int n, int c0, int c1, const std::vector<int32_t>& g2l) {
    const double h = 1.0 / n, a = 0.5 - 0.5 / std::sqrt(3.0), b = 0.5 + 0.5 / std::sqrt(3.0);
    const double xq[2] = {a, b};
    qw = {0.5, 0.5};
    for (int q = 0; q < 2; ++q) {
      val.push_back(1 - xq[q]); val.push_back(xq[q]);            // [nvar=1][nq][ndofs]
      grad.push_back(-1.0); grad.push_back(1.0);                 // [1][nq][ndofs][dref=1]
      geo_grad.push_back(-1.0); geo_grad.push_back(1.0);         // [1][nq][nv][1]
    }
    for (int c = c0; c < c1; ++c) {
      coords.push_back(c * h); coords.push_back((c + 1) * h);
      dofs.push_back(g2l.empty() ? c : g2l[c]); dofs.push_back(g2l.empty() ? c + 1 : g2l[c + 1]);
    }
  }
  vlfs_domain_desc desc(int ncells) {
    vlfs_domain_desc d;
    memset(&d, 0, sizeof d);
    d.D = 1; d.nq = nq; d.nslots = 1; d.measure_slot = 0; d.measure_kind = VLFS_MEASURE_CELL; d.nweights = 0; d.affine = 1;
    d.ncells = ncells; d.qweights = qw.data();
    vlfs_slot_desc& s = d.slot[0];
    s.dref = 1; s.nv = nv; s.nvar = 1; s.ndofs = nd; s.coords = coords.data(); s.geo_grad = geo_grad.data();
    s.val = val.data(); s.grad = grad.data(); s.dofs = dofs.data();
    return d;
  }
};

static int add_forms(vlfs_ctx* ctx, int dom, double sigma) {
  vlfs_term_desc t;
  memset(&t, 0, sizeof t);
  t.mat = 0; t.weight = -1; t.coef_re = 1.0; t.test_op = VLFS_OP_GRAD; t.trial_op = VLFS_OP_GRAD;
  t.test_scale[0] = 1.0; t.trial_scale[0] = 1.0;
  CHECK(vlfs_add_term(ctx, dom, &t, nullptr));
  t.coef_re = sigma; t.test_op = VLFS_OP_VAL; t.trial_op = VLFS_OP_VAL;
  CHECK(vlfs_add_term(ctx, dom, &t, nullptr));
  vlfs_rhs_desc r;
  memset(&r, 0, sizeof r);
  r.vec = 0; r.weight_re = r.weight_im = -1; r.coef_re = 1.0; r.op = VLFS_OP_VAL; r.scale[0] = 1.0;
  CHECK(vlfs_add_rhs_term(ctx, dom, &r, nullptr));
  return 0;
}

// analytic entries of K + sigma M for P1 on a uniform mesh
static void reference(int n, double sigma, std::vector<double>& diag, std::vector<double>& off, std::vector<double>& rhs) {
  const double h = 1.0 / n;
  diag.assign(n + 1, 0); off.assign(n, -1.0 / h + sigma * h / 6); rhs.assign(n + 1, 0);
  for (int c = 0; c < n; ++c)
    for (int k = 0; k < 2; ++k) { diag[c + k] += 1.0 / h + sigma * h / 3; rhs[c + k] += h / 2; }
}

static int run_single(int n, double sigma, std::vector<double>& x_out) {
  vlfs_ctx* ctx = nullptr;
  int rc = vlfs_create(&ctx, 0);
  if (rc != VLFS_OK) return 77;
  CHECK(vlfs_system_init(ctx, n + 1, 0, 1, 1));
  Mesh1D mesh(n, 0, n, {});
  vlfs_domain_desc d = mesh.desc(n);
  int dom = -1;
  CHECK(vlfs_add_domain(ctx, &d, &dom));
  if (add_forms(ctx, dom, sigma)) return 1;
  int64_t nnz = 0, ncoo = 0;
  CHECK(vlfs_build_pattern(ctx, &nnz, &ncoo));
  if (nnz != 3 * (n + 1) - 2 || ncoo != 4 * n) { fprintf(stderr, "pattern: nnz %lld ncoo %lld\n", (long long)nnz, (long long)ncoo); return 1; }
  CHECK(vlfs_assemble(ctx));
  std::vector<int32_t> rp(n + 2), ci(nnz);
  std::vector<double> v(nnz), b(n + 1), diag, off, rhs;
  CHECK(vlfs_get_pattern(ctx, rp.data(), ci.data()));
  CHECK(vlfs_get_values(ctx, 0, v.data()));
  CHECK(vlfs_get_vector(ctx, 0, b.data()));
  reference(n, sigma, diag, off, rhs);
  double err = 0;
  for (int i = 0; i <= n; ++i) {
    for (int p = rp[i]; p < rp[i + 1]; ++p) {
      const int j = ci[p];
      const double ref = j == i ? diag[i] : off[j < i ? j : i];
      if (std::abs(j - i) > 1) { fprintf(stderr, "unexpected entry (%d,%d)\n", i, j); return 1; }
      err = std::fmax(err, std::fabs(v[p] - ref) / std::fabs(diag[1]));
    }
    err = std::fmax(err, std::fabs(b[i] - rhs[i]) / rhs[1]);
  }
  if (err > 1e-13) { fprintf(stderr, "entries differ: %.3e\n", err); return 1; }
  // y = A x on host buffers
  std::vector<double> x(n + 1), y(n + 1);
  for (int i = 0; i <= n; ++i) x[i] = std::sin(0.1 * i);
  CHECK(vlfs_spmv(ctx, 0, x.data(), y.data()));
  for (int i = 0; i <= n; ++i) {
    double ref = diag[i] * x[i] + (i > 0 ? off[i - 1] * x[i - 1] : 0) + (i < n ? off[i] * x[i + 1] : 0);
    if (std::fabs(y[i] - ref) > 1e-12 * std::fabs(diag[1])) { fprintf(stderr, "spmv row %d\n", i); return 1; }
  }
  // a work matrix as the target of vlfs_combine: W = 2.5 A, multiplied through the same entry point
  {
    int wm = -1;
    const int src0 = 0;
    const double c25[2] = {2.5, 0.0};
    CHECK(vlfs_add_work_matrix(ctx, &wm));
    if (wm != 1) { fprintf(stderr, "work matrix index %d\n", wm); return 1; }
    CHECK(vlfs_combine(ctx, wm, 1, &src0, c25));
    std::vector<double> yw(n + 1);
    CHECK(vlfs_spmv(ctx, wm, x.data(), yw.data()));
    for (int i = 0; i <= n; ++i)
      if (std::fabs(yw[i] - 2.5 * y[i]) > 1e-12 * std::fabs(diag[1])) { fprintf(stderr, "work matrix row %d\n", i); return 1; }
    vlfs_term_desc bad; memset(&bad, 0, sizeof bad); bad.mat = wm;
    if (vlfs_add_term(ctx, 0, &bad, nullptr) == VLFS_OK) { fprintf(stderr, "term on a work matrix accepted\n"); return 1; }
  }
  std::vector<double> xyz(n + 1);
  for (int i = 0; i <= n; ++i) xyz[i] = (double)i / n;
  CHECK(vlfs_set_dof_coords(ctx, 1, xyz.data()));
  vlfs_solver_opts o;
  memset(&o, 0, sizeof o);
  o.method = VLFS_SOLVER_GMRES; o.precond = VLFS_PRECOND_SPARSE_LU; o.restart = 10; o.maxit = 20; o.rtol = 1e-12;
  CHECK(vlfs_solver_setup(ctx, 0, &o));
  vlfs_solve_stats st;
  x_out.assign(n + 1, 0.0);
  CHECK(vlfs_solve_vec(ctx, 0, x_out.data(), &st));
  if (!st.converged || st.rel_residual > 1e-12) { fprintf(stderr, "solve: rel %.3e\n", st.rel_residual); return 1; }
  // state errors come back as statuses with a message
  if (vlfs_build_pattern(ctx, &nnz, &ncoo) != VLFS_ERR_STATE || strlen(vlfs_last_error(ctx)) == 0) { fprintf(stderr, "second build not refused\n"); return 1; }
  // (2) the LinearSolver drop-in: the same matrix imported as CSR, no coordinates
  vlfs_ctx* c2 = nullptr;
  if (vlfs_create(&c2, 0) != VLFS_OK) return 1;
  { vlfs_ctx* ctx = c2;
    CHECK(vlfs_import_csr(ctx, n + 1, 0, 1, rp.data(), ci.data()));
    CHECK(vlfs_set_values(ctx, 0, v.data()));
    CHECK(vlfs_solver_setup(ctx, 0, &o));
    std::vector<double> x2(n + 1, 0.0);
    CHECK(vlfs_solve(ctx, b.data(), x2.data(), &st));
    double e = 0, nrm = 0;
    for (int i = 0; i <= n; ++i) { e += (x2[i] - x_out[i]) * (x2[i] - x_out[i]); nrm += x_out[i] * x_out[i]; }
    if (std::sqrt(e / nrm) > 1e-11) { fprintf(stderr, "imported solve differs: %.3e\n", std::sqrt(e / nrm)); return 1; }
    int64_t pert = -1, bad = -1;
    CHECK(vlfs_direct_status(ctx, &pert, &bad));
    if (pert != 0 || bad != 0) { fprintf(stderr, "pivot status %lld %lld\n", (long long)pert, (long long)bad); return 1; }
  }
  vlfs_destroy(c2);
  vlfs_destroy(ctx);
  return 0;
}

// the same chain on two ranks sharing device 0: rank 0 owns DOFs [0, m), rank 1 owns [m, n]
static int run_two_ranks(int n, double sigma, const std::vector<double>& x_ref) {
  vlfs_ctx* cs[2] = {nullptr, nullptr};
  const int devs[2] = {0, 0};
  vlfs_ctx* ctx = nullptr;
  int rc = vlfs_multi_create(cs, 2, devs);
  if (rc != VLFS_OK) { fprintf(stderr, "vlfs_multi_create -> %d\n", rc); return 1; }
  const int m = n / 2;                       // cut on the cell face at DOF m: it belongs to the higher rank
  // rank 0: owned 0..m-1, ghost m; cells 0..m-1.  rank 1: owned m..n, ghost m-1; cells m-1..n-1
  const int n_owned[2] = {m, n + 1 - m}, n_local[2] = {m + 1, n + 2 - m};
  for (int r = 0; r < 2; ++r) {
    ctx = cs[r];
    CHECK(vlfs_system_init(ctx, n_local[r], 0, 1, 1));
    std::vector<int32_t> g2l(n + 1, -1);
    if (r == 0) { for (int i = 0; i < m; ++i) g2l[i] = i; g2l[m] = m; }
    else { for (int i = m; i <= n; ++i) g2l[i] = i - m; g2l[m - 1] = n + 1 - m; }
    Mesh1D mesh(n, r == 0 ? 0 : m - 1, r == 0 ? m : n, g2l);
    vlfs_domain_desc d = mesh.desc(r == 0 ? m : n - m + 1);
    int dom = -1;
    CHECK(vlfs_add_domain(ctx, &d, &dom));
    if (add_forms(ctx, dom, sigma)) return 1;
    std::vector<double> xyz(n_local[r]);
    for (int i = 0; i <= n; ++i) if (g2l[i] >= 0) xyz[g2l[i]] = (double)i / n;
    CHECK(vlfs_set_dof_coords(ctx, 1, xyz.data()));
  }
  ctx = cs[0];
  CHECK(vlfs_multi_build_pattern(cs, 2, nullptr, nullptr));
  {  // halo plans: rank 0 sends its last owned DOF to rank 1 and receives DOF m; rank 1 the reverse
    const int32_t peer0[1] = {1}, peer1[1] = {0};
    const int64_t sp[2] = {0, 1}, rp[2] = {0, 1};
    const int32_t send0[1] = {m - 1}, send1[1] = {0};
    ctx = cs[0]; CHECK(vlfs_dist_set_halo(ctx, n_owned[0], 1, peer0, sp, send0, rp));
    ctx = cs[1]; CHECK(vlfs_dist_set_halo(ctx, n_owned[1], 1, peer1, sp, send1, rp));
  }
  ctx = cs[0];
  CHECK(vlfs_multi_assemble(cs, 2));
  vlfs_solver_opts o;
  memset(&o, 0, sizeof o);
  o.method = VLFS_SOLVER_GMRES; o.precond = VLFS_PRECOND_SPARSE_LU; o.restart = 10; o.maxit = 20; o.rtol = 1e-12;
  CHECK(vlfs_multi_solver_setup(cs, 2, 0, &o));
  std::vector<double> x0(n_local[0], 0.0), x1(n_local[1], 0.0);
  void* xs[2] = {x0.data(), x1.data()};
  vlfs_solve_stats st[2];
  rc = vlfs_multi_solve_vec(cs, 2, 0, xs, st);
  if (rc != VLFS_OK || !st[0].converged) { fprintf(stderr, "distributed solve: status %d rel %.3e (%s | %s)\n", rc, st[0].rel_residual, vlfs_last_error(cs[0]), vlfs_last_error(cs[1])); return 1; }
  double e = 0, nrm = 0;
  for (int i = 0; i <= n; ++i) {
    const double xi = i < m ? x0[i] : x1[i - m];
    e += (xi - x_ref[i]) * (xi - x_ref[i]); nrm += x_ref[i] * x_ref[i];
  }
  if (std::sqrt(e / nrm) > 1e-10) { fprintf(stderr, "two ranks differ from one: %.3e\n", std::sqrt(e / nrm)); return 1; }
  double info[4];
  ctx = cs[1];
  CHECK(vlfs_dist_info(ctx, info));
  if (info[0] != 1 || info[1] != 0) { fprintf(stderr, "cut sizes %g %g\n", info[0], info[1]); return 1; }
  vlfs_destroy(cs[0]); vlfs_destroy(cs[1]);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "--layout")) return layout();
  if (vlfs_version() < 100) return 1;
  const int n = 64;
  std::vector<double> x;
  int rc = run_single(n, 2.5, x);
  if (rc) return rc;
  rc = run_two_ranks(n, 2.5, x);
  if (rc) return rc;
  printf("cabi_harness: all checks passed\n");
  return 0;
}
