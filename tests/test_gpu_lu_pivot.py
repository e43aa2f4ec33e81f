"""Hardening of the GPU multifrontal LU (the `LUSolver()` = UMFPACK twin, src/Periodic_Beam.jl:107,
src/Yago_freq_domain.jl:171): threshold partial pivoting inside the pivot blocks, perturbation and
counting of tiny pivots, detection of non-finite pivots, and the solver as a drop-in on a matrix
the host already holds (`numerical_setup(ss, A)`), with and without DOF coordinates."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


def _solve(A, b, coords=None, rtol=1e-12):
    import vlfs_b200
    from vlfs_b200 import _lib as L
    S = vlfs_b200.B200System.from_csr(A)
    if coords is not None:
        S.set_dof_coords(coords)
    S.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=10, maxit=20, rtol=rtol)
    x, st = S.solve(b=b)
    return S, x, st


def _grid_laplacian(n, complex_shift=False):
    """5-point Laplacian on an n x n grid with lexicographic numbering + coordinates."""
    T = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(n, n))
    A = (sp.kron(sp.eye(n), T) + sp.kron(T, sp.eye(n))).tocsr()
    if complex_shift:
        A = (A - (0.3 + 0.2j) * sp.eye(n * n)).tocsr()
    ij = np.stack(np.meshgrid(np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 2).astype(float)
    return A, ij


@pytest.mark.parametrize("cplx", [False, True])
def test_imported_matrix_solves_with_and_without_coordinates(cplx):
    A, xy = _grid_laplacian(40, complex_shift=cplx)
    rng = np.random.default_rng(3)
    b = rng.standard_normal(A.shape[0]) + (1j * rng.standard_normal(A.shape[0]) if cplx else 0)
    xr = spla.splu(A.tocsc()).solve(b)
    for coords in (xy, None):                 # None: level-structure dissection on graph distances
        S, x, st = _solve(A, b, coords)
        assert st["converged"], st
        assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-10
        assert S.direct_status() == (0, 0)


def test_zero_diagonal_needs_the_in_block_pivoting():
    """Saddle-point-like matrix with structurally zero diagonal entries: an LU without row
    interchanges divides by zero; the in-block threshold pivoting solves it to LU accuracy."""
    n = 24
    rng = np.random.default_rng(7)
    B = rng.standard_normal((n, n))
    A = B + B.T
    np.fill_diagonal(A, 0.0)                  # every leading pivot is exactly zero
    A[np.arange(0, n, 2), np.arange(1, n, 2)] += 4.0
    A[np.arange(1, n, 2), np.arange(0, n, 2)] += 4.0
    A = sp.csr_matrix(A)
    b = rng.standard_normal(n)
    S, x, st = _solve(A, b)
    xr = np.linalg.solve(A.toarray(), b)
    assert st["converged"], st
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-10
    pert, bad = S.direct_status()
    assert bad == 0 and pert == 0             # interchanges, not perturbations, did the job


def test_sparse_matrix_with_structurally_zero_diagonal():
    """Sparse system of 900 two-unknown nodes on a 30 x 30 grid whose diagonal is exactly zero (each
    node is a [[0,3],[3,0]] block, neighbours couple weakly): every front meets zero pivots and the
    row interchanges inside the pivot blocks must resolve them - solution to SuperLU accuracy,
    nothing perturbed."""
    m = 30
    rng = np.random.default_rng(11)
    G, xy = _grid_laplacian(m)
    G = G.tocoo()
    rows, cols, vals = [], [], []
    for i, j in zip(G.row, G.col):
        if i == j:
            blk = np.array([[0.0, 3.0], [3.0, 0.0]])
        else:
            blk = 0.2 * rng.standard_normal((2, 2))
        for a_ in range(2):
            for b_ in range(2):
                if i == j and a_ == b_:
                    continue
                rows.append(2 * i + a_); cols.append(2 * j + b_); vals.append(blk[a_, b_])
    n = 2 * m * m
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    A.sort_indices()
    assert np.all(A.diagonal() == 0.0)
    b = rng.standard_normal(n)
    xr = spla.splu(A.tocsc()).solve(b)
    S, x, st = _solve(A, b, np.repeat(xy, 2, axis=0), rtol=1e-11)
    assert st["converged"], (st, S.direct_status())
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-9
    assert S.direct_status() == (0, 0)


def test_singular_matrix_reports_a_status_not_garbage():
    import vlfs_b200
    from vlfs_b200 import _lib as L
    n = 16
    A = sp.csr_matrix(np.ones((n, n)))        # rank one: every Schur complement vanishes
    S = vlfs_b200.B200System.from_csr(A)
    o = L.SolverOpts(L.SOLVER_GMRES, L.PRECOND_SPARSE_LU, 5, 5, 1e-10, 0, 0)
    import ctypes as C
    rc = S.lib.vlfs_solver_setup(S.ctx, 0, C.byref(o))
    assert rc in (L.PIVOT_PERTURBED, L.ERR_SINGULAR), rc
    pert, bad = S.direct_status()
    assert pert + bad > 0
    assert b"pivot" in S.lib.vlfs_last_error(S.ctx)
    A2 = sp.csr_matrix(np.full((n, n), np.nan))
    S2 = vlfs_b200.B200System.from_csr(A2)
    rc2 = S2.lib.vlfs_solver_setup(S2.ctx, 0, C.byref(o))
    assert rc2 == L.ERR_SINGULAR


def test_second_pattern_build_is_refused_and_refactor_repoints_the_residual_matrix():
    import vlfs_b200
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    p = YagoProblem(Yago_freq_domain_params(nx=2, ny=1, nz=1, order=2, lfactor=0.4, dfactor=3), device=0)
    p.assemble()
    with pytest.raises(L.VlfsError) as e:
        p.sys.build_pattern()
    assert "already built" in str(e.value)
    # two value arrays on one pattern: refactor(1) must make the Krylov residual use matrix 1
    A = p.sys.get_matrix(0)
    S = vlfs_b200.B200System.from_csr(A, nmat=2)
    S.set_values(A.data, 0)
    S.set_values(2.0 * A.data, 1)
    S.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=10, maxit=20, rtol=1e-11)
    b = p.sys.get_vector(0)
    x0, st0 = S.solve(b=b)
    S.refactor(1)
    x1, st1 = S.solve(b=b)
    assert st0["converged"] and st1["converged"]
    assert np.linalg.norm(2.0 * x1 - x0) / np.linalg.norm(x0) < 1e-9


def test_newmark_rejects_a_solver_set_up_on_another_operator():
    import vlfs_b200
    from vlfs_b200 import _lib as L
    A, xy = _grid_laplacian(12)
    n = A.shape[0]
    S = vlfs_b200.B200System.from_csr(A, nmat=4)
    K, Cm, M = A.data, 0.1 * A.data, np.where(A.indices == np.repeat(np.arange(n), np.diff(A.indptr)), 1.0, 0.0)
    for k, v in enumerate((K, Cm, M)):
        S.set_values(v, k)
    dt, gamma, beta = 0.01, 0.5, 0.25
    S.combine(3, [0, 1, 2], [1.0, gamma / (beta * dt), 1.0 / (beta * dt * dt)])
    S.set_dof_coords(xy)
    S.solver_setup(3, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=5, maxit=10, rtol=1e-12)
    u0 = np.sin(np.arange(n) * 0.3)
    z = np.zeros(n)
    u, v, a, out, st = S.newmark_run(0, 1, 2, 20, dt, gamma, beta, None, u0, z, z)
    assert st["converged"] and np.isfinite(u).all()
    # reference: the same Newmark recurrences with SciPy
    Km, Cmm, Mm = (sp.csr_matrix((d, A.indices, A.indptr), shape=A.shape) for d in (K, Cm, M))
    cu, cuu = gamma / (beta * dt), 1.0 / (beta * dt * dt)
    lu = spla.splu((Km + cu * Cmm + cuu * Mm).tocsc())
    ur, vr, ar = u0.copy(), z.copy(), z.copy()
    for _ in range(20):
        vs = -cu * ur + (1 - gamma / beta) * vr + dt * (1 - gamma / (2 * beta)) * ar
        as_ = -cuu * ur - vr / (beta * dt) - (1 - 2 * beta) / (2 * beta) * ar
        un = lu.solve(-Cmm @ vs - Mm @ as_)
        ur, vr, ar = un, vs + cu * un, as_ + cuu * un
    assert np.linalg.norm(u - ur) / np.linalg.norm(ur) < 1e-9
    with pytest.raises(L.VlfsError) as e:       # other dt: the factorised operator no longer matches
        S.newmark_run(0, 1, 2, 5, 2 * dt, gamma, beta, None, u0, z, z)
    assert "not K + gamma" in str(e.value)
