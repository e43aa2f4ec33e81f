"""Error behaviour and edge inputs of the C ABI on the GPU: every misuse returns a status and a
message through `vlfs_last_error` (no abort, no exception across the boundary),
empty triangulations are legal, and a system that is used again after an error still works."""
import ctypes as C
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KW = dict(nx=2, ny=1, nz=1, order=2, lfactor=0.4, dfactor=3)


@pytest.fixture(scope="module")
def prob():
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    p = YagoProblem(Yago_freq_domain_params(**KW), device=0)
    p.assemble()
    return p


def test_calls_out_of_order_report_state_errors():
    import vlfs_b200
    from vlfs_b200 import _lib as L
    S = vlfs_b200.B200System(10, False)
    buf = np.zeros(16)
    for call in (S.assemble, S.build_pattern, lambda: S.solver_setup(0), lambda: S.refactor(0),
                 lambda: S._check(S.lib.vlfs_get_values(S.ctx, 0, buf.ctypes.data_as(C.c_void_p)))):
        with pytest.raises(L.VlfsError) as e:
            call()
        assert "status -" in str(e.value)                      # argument / state error, not a CUDA error


def test_bad_indices_are_argument_errors_and_leave_the_system_usable(prob):
    from vlfs_b200 import _lib as L
    S = prob.sys
    v0 = S.get_values(0).copy()
    bad = [lambda: S._check(S.lib.vlfs_get_values(S.ctx, 7, v0.ctypes.data_as(C.c_void_p))),
           lambda: S._check(S.lib.vlfs_get_vector(S.ctx, -1, v0.ctypes.data_as(C.c_void_p))),
           lambda: S._check(S.lib.vlfs_set_term_coef(S.ctx, 99, 0, 1.0, 0.0)),
           lambda: S._check(S.lib.vlfs_set_weights(S.ctx, 1, 99, v0.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)))),
           lambda: S.combine(9, [0], [1.0]),
           lambda: S.solver_setup(5),
           lambda: S.refactor(0)]                              # no sparse LU set up yet
    for call in bad:
        with pytest.raises(L.VlfsError):
            call()
    msg = S.lib.vlfs_last_error(S.ctx)
    assert msg and len(msg.decode()) > 0
    prob.assemble()                                            # still usable, same result
    assert np.array_equal(S.get_values(0).view(np.uint64), v0.view(np.uint64))


def test_nonconvergence_is_a_status_not_an_error(prob):
    """Jacobi-GMRES cannot solve the indefinite monolithic system in 5 iterations: the call returns
    VLFS_NOT_CONVERGED with statistics ('status + stats, not an error')."""
    from vlfs_b200 import _lib as L
    S = prob.sys
    S.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_JACOBI, restart=5, maxit=5, rtol=1e-12)
    x, st = S.solve(vec=0)
    assert not st["converged"] and st["iterations"] == 5 and st["rel_residual"] > 1e-12
    S.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=10, maxit=20, rtol=1e-12)
    x, st = S.solve(vec=0)
    r = S.spmv(x) - S.get_vector(0)
    assert np.linalg.norm(r) <= 1e-9 * np.linalg.norm(S.get_vector(0))


def test_empty_triangulation_is_legal():
    """A domain without cells (a rank that holds no plate cells, a tag without facets) contributes
    nothing and breaks nothing."""
    import vlfs_b200
    from vlfs_b200 import _lib as L
    from vlfs_b200.geometry import CartesianDiscreteModel, add_tag_from_tags, Interior, Boundary, Triangulation
    from vlfs_b200.fe import ReferenceFE, TestFESpace, MultiFieldFESpace, TrialFESpace, Measure, lagrangian
    model = CartesianDiscreteModel((0, 1, 0, 1), (3, 2))
    add_tag_from_tags(model, "top", [3, 4, 6])                   # top corners and top side
    Om = Interior(model)
    G = Boundary(model, tags="top")
    Ge = Triangulation(G, np.zeros(0, dtype=np.int64))         # no facets at all
    V = TestFESpace(Om, ReferenceFE(lagrangian, float, 2))
    X = MultiFieldFESpace([TrialFESpace(V)])
    dOm, dG, dGe = Measure(Om, 4), Measure(G, 4), Measure(Ge, 4)
    aux = []                                                    # surface slots that carry the measure
    for t in (G, Ge):
        Va = TestFESpace(t, ReferenceFE(lagrangian, float, 1))
        Va.offset = 0
        aux.append(Va)
    results = []
    for with_empty in (False, True):
        S = vlfs_b200.B200System(X.ndofs, False, nmat=1, nvec=1)
        d = S.add_domain(2, dOm.nq, dOm.wq, [dOm.own_slot(V)])
        d.add_term(0, 1.0, L.OP_GRAD, 0, L.OP_GRAD, 0)
        d.add_term(0, 2.0, L.OP_VAL, 0, L.OP_VAL, 0)
        d = S.add_domain(2, dG.nq, dG.wq, [dG.trace_slot(V), dG.surface_slot(aux[0])], measure_slot=1)
        d.add_term(0, 3.0, L.OP_VAL, 0, L.OP_VAL, 0)
        d.add_rhs(0, 1.0, L.OP_VAL, 0)
        if with_empty:
            d = S.add_domain(2, dGe.nq, dGe.wq, [dGe.trace_slot(V), dGe.surface_slot(aux[1])], measure_slot=1)
            d.add_term(0, 5.0, L.OP_VAL, 0, L.OP_VAL, 0)
            d.add_rhs(0, 7.0, L.OP_VAL, 0)
        S.build_pattern()
        S.assemble()
        results.append((S.get_pattern(), S.get_values(0), S.get_vector(0)))
    (p0, v0, b0), (p1, v1, b1) = results
    assert np.array_equal(p0[0], p1[0]) and np.array_equal(p0[1], p1[1])
    assert np.array_equal(v0, v1) and np.array_equal(b0, b1)
    # the matrix is the SPD operator ∫∇u·∇v + 2∫uv + 3∫_Γ uv: symmetric, positive row sums of the mass parts
    import scipy.sparse as sp
    A = sp.csr_matrix((v0, p0[1], p0[0]), shape=(X.ndofs, X.ndofs))
    assert abs(A - A.T).max() <= 1e-14 * abs(A).max()
    assert abs(A.sum() - (2.0 * 1.0 + 3.0 * 1.0)) <= 1e-12      # ∑ of all entries = 2|Ω| + 3|Γ|
    assert abs(b0.sum() - 1.0) <= 1e-13                         # ∫_Γ 1 = |Γ| = 1


def test_levels_longer_than_the_grid_limit_are_split(prob, monkeypatch):
    """The LU launches are batched over the fronts of a tree level through gridDim.y/z (<= 65 535);
    longer levels are processed in consecutive pieces.  Forced here with a limit of 3 fronts (the
    analysis is made once per context, so the second system is a fresh one)."""
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    opts = dict(method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=10, maxit=20, rtol=1e-13)
    S = prob.sys
    S.solver_setup(0, **opts)
    x0, st0 = S.solve(vec=0)
    info0 = S.direct_info()
    monkeypatch.setenv("VLFS_LU_ZMAX", "3")
    p2 = YagoProblem(Yago_freq_domain_params(**KW), device=0)
    p2.assemble()
    p2.sys.solver_setup(0, **opts)
    x1, st1 = p2.sys.solve(vec=0)
    info1 = p2.sys.direct_info()
    assert info1["nfronts"] == info0["nfronts"] and info1["nlevels"] > info0["nlevels"]
    assert np.linalg.norm(x1 - x0) <= 1e-11 * np.linalg.norm(x0)
