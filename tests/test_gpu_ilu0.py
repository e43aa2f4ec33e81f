"""Block ILU(0) preconditioner (north_star (5): "CG / GMRES / BiCGStab with Jacobi or block-ILU0"), `ilu0.cu`
through the C ABI against the sequential restatement `oracle/ilu0.py`: factors applied to vectors, chains and
wide dependency levels, real and complex, the owned block of a partitioned context, refactorisation, and
Krylov convergence on the Newmark matrix of configuration 5-1-1."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


def _system(A, precond, method=None, **kw):
    import vlfs_b200
    from vlfs_b200 import _lib as L
    S = vlfs_b200.B200System.from_csr(A)
    opts = dict(method=L.SOLVER_GMRES if method is None else method, precond=precond, restart=60, maxit=400, rtol=1e-11)
    opts.update(kw)
    return S, opts


def _grid(n, cplx):
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    A = (sp.kron(sp.eye(n), T) + sp.kron(T, sp.eye(n)) + 0.05 * sp.eye(n * n)).tocsr()
    if cplx:
        A = (A + 0.3j * sp.kron(sp.eye(n), T)).tocsr()
    return A


def _fe_like(cplx, seed=0):
    """Pattern and values of a real FE matrix (Yago warm-up, 9 fields coupled), made diagonally dominant so
    that the incomplete factors are well conditioned and the comparison is sharp."""
    from oracle.cases.yago import YagoCase
    A, _ = YagoCase(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3).assemble()
    A = A.tocsr(); A.sort_indices()
    P = A.copy(); P.data = np.ones_like(P.data)
    rng = np.random.default_rng(seed)
    vals = rng.standard_normal(P.nnz) + (1j * rng.standard_normal(P.nnz) if cplx else 0)
    B = sp.csr_matrix((vals, P.indices.copy(), P.indptr.copy()), shape=P.shape)
    rowsum = np.asarray(abs(B).sum(axis=1)).ravel()
    return (B + sp.diags(1.5 * rowsum)).tocsr()


@pytest.mark.parametrize("cplx", [False, True], ids=["real", "complex"])
def test_tridiagonal_chain_is_the_exact_lu(cplx):
    """No fill on a tridiagonal matrix: ILU(0) = LU, one dependency chain of n levels (the single-CTA walk)."""
    from vlfs_b200 import _lib as L
    n = 3000
    rng = np.random.default_rng(1)
    lo, up = rng.standard_normal(n - 1), rng.standard_normal(n - 1)
    dg = 3.0 + rng.random(n) + (0.5j * rng.random(n) if cplx else 0)
    A = sp.diags([lo, dg, up], [-1, 0, 1]).tocsr()
    S, opts = _system(A, L.PRECOND_BLOCK_ILU0)
    S.solver_setup(0, **opts)
    info = S.precond_info()
    assert info["levels_lower"] == n and info["levels_upper"] == n and info["launches_lower"] == 1
    x = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    z = S.precond_apply(A @ x)
    assert np.abs(z - x).max() <= 1e-12 * np.abs(x).max()
    xs, st = S.solve(b=A @ x)
    assert st["converged"] and st["iterations"] <= 2
    assert np.linalg.norm(xs - x) <= 1e-10 * np.linalg.norm(x)


@pytest.mark.parametrize("cplx", [False, True], ids=["real", "complex"])
def test_factors_match_the_sequential_restatement_on_a_grid(cplx):
    """5-point grid: anti-diagonal levels, narrow at both ends (chains) and wide in the middle (one warp per row)."""
    from vlfs_b200 import _lib as L
    from oracle import ilu0 as O
    n = 70
    A = _grid(n, cplx)
    Lm, Um = O.ilu0(A)
    ll, ul = O.levels(A)
    S, opts = _system(A, L.PRECOND_BLOCK_ILU0)
    S.solver_setup(0, **opts)
    info = S.precond_info()
    assert info["levels_lower"] == ll.max() + 1 and info["levels_upper"] == ul.max() + 1
    assert info["launches_lower"] < info["levels_lower"]            # the narrow levels share launches
    rng = np.random.default_rng(2)
    for _ in range(2):
        r = rng.standard_normal(n * n) + (1j * rng.standard_normal(n * n) if cplx else 0)
        z, zr = S.precond_apply(r), O.apply(Lm, Um, r)
        assert np.abs(z - zr).max() <= 1e-12 * np.abs(zr).max()
    assert S.direct_status() == (0, 0)


@pytest.mark.parametrize("cplx", [False, True], ids=["real", "complex"])
def test_factors_match_on_a_finite_element_pattern(cplx):
    """~60 entries per row with irregular dependencies (the Yago warm-up pattern): every a_ij receives its
    updates in ascending k as in the sequential algorithm, so the factors agree to rounding; a second set
    of values goes through `vlfs_solver_refactor` on the same levels."""
    from vlfs_b200 import _lib as L
    from oracle import ilu0 as O
    A = _fe_like(cplx)
    S, opts = _system(A, L.PRECOND_BLOCK_ILU0)
    S.solver_setup(0, **opts)
    rng = np.random.default_rng(3)
    r = rng.standard_normal(A.shape[0]) + (1j * rng.standard_normal(A.shape[0]) if cplx else 0)
    Lm, Um = O.ilu0(A)
    zr = O.apply(Lm, Um, r)
    z = S.precond_apply(r)
    assert np.abs(z - zr).max() <= 1e-11 * np.abs(zr).max()
    z2 = S.precond_apply(r)
    assert np.array_equal(z, z2)                                     # bitwise repeatable
    A2 = _fe_like(cplx, seed=5)
    S.set_values(A2.data, 0)
    S.refactor(0)
    Lm, Um = O.ilu0(A2)
    z, zr = S.precond_apply(r), O.apply(Lm, Um, r)
    assert np.abs(z - zr).max() <= 1e-11 * np.abs(zr).max()
    xs, st = S.solve(b=A2 @ r)
    assert st["converged"] and np.linalg.norm(xs - r) <= 1e-9 * np.linalg.norm(r)


def test_block_of_the_owned_rows():
    """On a partitioned context the preconditioner is the ILU(0) of owned rows x owned columns: ghost
    columns are dropped (block Jacobi across GPUs, ILU(0) inside)."""
    from vlfs_b200 import _lib as L
    from oracle import ilu0 as O
    A = _grid(40, True)
    n_owned = 1000
    S, opts = _system(A, L.PRECOND_BLOCK_ILU0)
    S.set_owned_rows(n_owned)
    S.solver_setup(0, **opts)
    Lm, Um = O.ilu0(A, n_owned)
    rng = np.random.default_rng(4)
    r = rng.standard_normal(A.shape[0]) + 1j * rng.standard_normal(A.shape[0])
    z = S.precond_apply(r)[:n_owned]
    zr = O.apply(Lm, Um, r[:n_owned])
    assert np.abs(z - zr).max() <= 1e-12 * np.abs(zr).max()


def test_krylov_with_ilu0_on_the_newmark_matrix():
    """Configuration 5-1-1 (src/Periodic_Beam.jl:104-108): A = K + γ/(βΔt) C + 1/(βΔt²) M assembled on the GPU;
    GMRES and BiCGStab preconditioned by ILU(0) reach the LU solution (≤ 1e-8) in fewer iterations than Jacobi."""
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.periodic_beam import PeriodicBeamProblem, Periodic_Beam_params, MAT_A
    prob = PeriodicBeamProblem(Periodic_Beam_params(n=8, dt=0.01, tf=0.02, orderphi=2, ordereta=2, k=3), device=0)
    prob.assemble()
    A = prob.sys.get_matrix(MAT_A).tocsr()
    rng = np.random.default_rng(6)
    xt = rng.standard_normal(A.shape[0])
    b = A @ xt
    xr = spla.splu(A.tocsc()).solve(b)
    its = {}
    for name, method, precond in [("gmres_ilu0", L.SOLVER_GMRES, L.PRECOND_BLOCK_ILU0),
                                  ("bicgstab_ilu0", L.SOLVER_BICGSTAB, L.PRECOND_BLOCK_ILU0),
                                  ("gmres_jacobi", L.SOLVER_GMRES, L.PRECOND_JACOBI)]:
        prob.sys.solver_setup(MAT_A, method=method, precond=precond, restart=100, maxit=3000, rtol=1e-11)
        x, st = prob.sys.solve(b=b)
        its[name] = st["iterations"]
        print(name, st["iterations"], st["rel_residual"], st["converged"])
        if "ilu0" in name:
            assert st["converged"], (name, st)
            assert np.linalg.norm(x - xr) <= 1e-8 * np.linalg.norm(xr), name
    assert its["gmres_ilu0"] < its["gmres_jacobi"]


def test_cg_gmres_bicgstab_with_jacobi_and_ilu0_on_an_spd_matrix():
    """north_star (5) on the matrix class CG is meant for (the surface mass / Laplace blocks are symmetric
    positive definite): every method x {none, Jacobi, ILU(0)} reaches the LU solution; ILU(0) needs the fewest
    iterations.  (ILU(0) of a symmetric matrix is symmetric: U = D Lᵀ, so it is admissible for CG.)"""
    from vlfs_b200 import _lib as L
    n = 48
    A = _grid(n, False)
    A = (sp.diags(1.0 + np.arange(n * n) % 7) @ A @ sp.diags(1.0 + np.arange(n * n) % 7)).tocsr()   # uneven diagonal
    rng = np.random.default_rng(8)
    xt = rng.standard_normal(n * n)
    b = A @ xt
    S, _ = _system(A, L.PRECOND_NONE)
    its = {}
    for mname, method in [("cg", L.SOLVER_CG), ("gmres", L.SOLVER_GMRES), ("bicgstab", L.SOLVER_BICGSTAB)]:
        for pname, precond in [("none", L.PRECOND_NONE), ("jacobi", L.PRECOND_JACOBI), ("ilu0", L.PRECOND_BLOCK_ILU0)]:
            S.solver_setup(0, method=method, precond=precond, restart=80, maxit=4000, rtol=1e-11)
            x, st = S.solve(b=b)
            its[mname, pname] = st["iterations"]
            print(mname, pname, st["iterations"], st["rel_residual"])
            if pname != "none":
                assert st["converged"], (mname, pname, st)
                assert np.linalg.norm(x - xt) <= 1e-8 * np.linalg.norm(xt), (mname, pname)
    for mname in ("cg", "gmres", "bicgstab"):
        assert its[mname, "ilu0"] < its[mname, "jacobi"], its
    # SciPy's CG on the same matrix with the restatement's factors: 295 / 117 / 35 iterations
    assert abs(its["cg", "none"] - 295) <= 3 and abs(its["cg", "jacobi"] - 117) <= 3 and abs(its["cg", "ilu0"] - 35) <= 2, its
    # GMRES minimises the residual over the same Krylov space: with an orthogonal basis it cannot need many more
    # iterations than CG while the restart length (80) is not reached
    assert its["gmres", "ilu0"] <= its["cg", "ilu0"] + 5, its


def _block_diagonal(A, B):
    A = A.tocsr()
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    keep = (rows // B) == (A.indices // B)
    return sp.csr_matrix((A.data[keep], (rows[keep], A.indices[keep])), shape=A.shape)


@pytest.mark.parametrize("cplx,B", [(False, 512), (True, 300)], ids=["real-512", "complex-300"])
def test_block_jacobi_of_ilu0_blocks(cplx, B):
    """`block_size_hint = B`: diagonal blocks of B consecutive rows, ILU(0) inside each, couplings between blocks
    dropped; one CTA per block walks the block's own levels.  Equals the restatement applied to the block-diagonal
    part of the matrix; as a preconditioner it sits between Jacobi and the ILU(0) of the whole matrix."""
    from vlfs_b200 import _lib as L
    from oracle import ilu0 as O
    n = 70
    A = _grid(n, cplx)
    Ab = _block_diagonal(A, B)
    Lm, Um = O.ilu0(Ab)
    ll, ul = O.levels(Ab)
    S, opts = _system(A, L.PRECOND_BLOCK_ILU0, block_size_hint=B)
    S.solver_setup(0, **opts)
    info = S.precond_info()
    assert info["levels_lower"] == ll.max() + 1 and info["levels_upper"] == ul.max() + 1
    assert info["launches_lower"] == 1 and info["launches_upper"] == 1
    rng = np.random.default_rng(9)
    r = rng.standard_normal(n * n) + (1j * rng.standard_normal(n * n) if cplx else 0)
    z, zr = S.precond_apply(r), O.apply(Lm, Um, r)
    assert np.abs(z - zr).max() <= 1e-12 * np.abs(zr).max()
    xt = rng.standard_normal(n * n) + (1j * rng.standard_normal(n * n) if cplx else 0)
    its = {}
    for name, kw in [("blocks", dict(precond=L.PRECOND_BLOCK_ILU0, block_size_hint=B)),
                     ("whole", dict(precond=L.PRECOND_BLOCK_ILU0, block_size_hint=0)),
                     ("jacobi", dict(precond=L.PRECOND_JACOBI))]:
        S.solver_setup(0, method=L.SOLVER_GMRES, restart=120, maxit=2000, rtol=1e-11, **kw)
        x, st = S.solve(b=A @ xt)
        assert st["converged"], (name, st)
        assert np.linalg.norm(x - xt) <= 1e-8 * np.linalg.norm(xt)
        its[name] = st["iterations"]
    print(its)
    assert its["whole"] <= its["blocks"] <= its["jacobi"], its
