"""GPU parity (through the C ABI) against the oracle for the remaining BASELINE.json
configurations: 5-1-1 periodic beam (time domain), 5-2-1 / 5-2-2 Khabakhpasheva (frequency and
time domain), 5-5-1 MultiGeo (tets).  Pattern bit-exact, entries <= 1e-12, solutions <= 1e-8."""
import os
import numpy as np
import pytest
import scipy.sparse.linalg as spla
from conftest import lu_reference

pytestmark = pytest.mark.gpu


def _csr(A):
    A = A.tocsr(); A.sort_indices()
    return A


def _check_matrix(sys_, mat, A, tol=1e-12):
    A = _csr(A)
    rp, ci = sys_.get_pattern()
    assert np.array_equal(rp, A.indptr) and np.array_equal(ci, A.indices)
    v = sys_.get_values(mat)
    assert np.abs(v - A.data).max() <= tol * max(np.abs(A.data).max(), 1e-300)
    return A


@pytest.mark.parametrize("kw", [dict(n=3, dt=0.5, tf=1.0, orderphi=2, ordereta=2, k=1),       # warm-up like
                                dict(n=8, dt=1e-6, tf=2e-5, orderphi=3, ordereta=3, k=15),
                                dict(n=4, dt=1e-6, tf=1e-5, orderphi=2, ordereta=4, k=15)])
def test_periodic_beam(kw):
    from vlfs_b200.cases.periodic_beam import PeriodicBeamProblem, Periodic_Beam_params, MAT_K, MAT_C, MAT_M, MAT_A
    from oracle.cases.periodic_beam import PeriodicBeamCase
    prob = PeriodicBeamProblem(Periodic_Beam_params(**kw), device=0)
    prob.assemble()
    ref = PeriodicBeamCase(**kw)
    K, C, M = ref.assemble()
    for mat, A in ((MAT_K, K), (MAT_C, C), (MAT_M, M)):
        _check_matrix(prob.sys, mat, A)
    cu, cuu = 0.5 / (0.25 * kw["dt"]), 1.0 / (0.25 * kw["dt"] ** 2)
    _check_matrix(prob.sys, MAT_A, K + cu * C + cuu * M)
    u, v, a, states, st = prob.solve()
    r = ref.run(keep_states=True)
    assert states.shape == (len(r["t"]), ref.X.ndofs)
    sref = np.array(r["states"])
    assert np.abs(states - sref).max() <= 1e-8 * np.abs(sref).max()
    assert np.abs(u - r["u"]).max() <= 1e-8 * np.abs(r["u"]).max()
    assert np.abs(v - r["v"]).max() <= 1e-7 * np.abs(r["v"]).max()
    # the reference's own acceptance quantity: L2 errors against the analytic wave
    l2_phi, l2_eta, _ = ref.functionals()
    t = r["t"][-1]
    assert abs(l2_phi(u, t) - r["e_phi"][-1]) <= 1e-6 * r["e_phi"][-1]
    assert abs(l2_eta(u, t) - r["e_eta"][-1]) <= 1e-6 * r["e_eta"][-1]


@pytest.mark.parametrize("kw", [dict(nx=10, ny=1, order=2, xi=0.0),           # scripts/5-2-1-...jl:17-28
                                dict(nx=10, ny=2, order=4, xi=625.0)])
def test_khabakhpasheva_freq(kw):
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.khabakhpasheva import KhabakhpashevaFreqProblem, Khabakhpasheva_freq_domain_params
    from oracle.cases.khabakhpasheva import KhabakhpashevaFreqCase
    prob = KhabakhpashevaFreqProblem(Khabakhpasheva_freq_domain_params(**kw), device=0)
    prob.assemble()
    ref = KhabakhpashevaFreqCase(**kw)
    A, b = ref.assemble()
    A = _check_matrix(prob.sys, 0, A)
    bb = prob.sys.get_vector(0)
    assert np.abs(bb - b).max() <= 1e-12 * np.abs(b).max()
    prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    x, st = prob.sys.solve(vec=0)
    xr = lu_reference(A, b)
    assert np.linalg.norm(x - xr) <= 1e-8 * np.linalg.norm(xr)
    xs, eta = prob.postprocess(x)
    xs_r, eta_r = ref.eta_postprocess(xr)
    assert np.array_equal(xs, xs_r) and np.abs(eta - eta_r).max() <= 1e-8 * eta_r.max()


@pytest.mark.parametrize("kw", [dict(nx=10, ny=1, order=2, xi=0.0), dict(nx=10, ny=2, order=4, xi=625.0)])
def test_khabakhpasheva_time(kw):
    from vlfs_b200.cases.khabakhpasheva import (KhabakhpashevaTimeProblem, Khabakhpasheva_time_domain_params,
                                                MAT_K, MAT_C, MAT_M)
    from oracle.cases.khabakhpasheva import KhabakhpashevaTimeCase
    prob = KhabakhpashevaTimeProblem(Khabakhpasheva_time_domain_params(**kw), device=0)
    prob.assemble()
    ref = KhabakhpashevaTimeCase(**kw)
    for mat, A in zip((MAT_K, MAT_C, MAT_M), ref.assemble()):
        _check_matrix(prob.sys, mat, A)
    nst = 60
    ts, hist, (u, v, a), st = prob.solve(nsteps=nst)
    ts_r, hist_r, (ur, vr, ar) = ref.run(nsteps=nst)
    assert np.allclose(ts, ts_r, rtol=1e-14, atol=0)
    assert np.abs(hist - hist_r).max() <= 1e-8 * np.abs(hist_r).max()
    assert np.abs(u - ur).max() <= 1e-8 * np.abs(ur).max()


@pytest.mark.parametrize("order", [2, 4])                                     # scripts/5-5-1-MultiGeo.jl:16-25
def test_multigeo_coarse(order, golden_dir):
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
    from oracle.cases.multigeo import MultiGeoCase
    f = os.path.join(golden_dir, "multi_geo_coarse.json")
    prob = MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=f, order=order), device=0)
    prob.assemble()
    ref = MultiGeoCase(f, order=order)
    A, b = ref.assemble()
    A = _check_matrix(prob.sys, 0, A, tol=2e-12)
    assert np.abs(prob.sys.get_vector(0) - b).max() <= 2e-12 * np.abs(b).max()
    prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    x, st = prob.sys.solve(vec=0)
    xr = lu_reference(A, b)
    assert np.linalg.norm(x - xr) <= 1e-8 * np.linalg.norm(xr)


def test_run_periodic_beam_functionals_on_gpu():
    """The drop-in `run_periodic_beam`: L2 errors and the four energies evaluated by the GPU
    functional kernel equal the oracle's (src/Periodic_Beam.jl:141-148)."""
    from vlfs_b200.cases.periodic_beam import run_periodic_beam, Periodic_Beam_params
    from oracle.cases.periodic_beam import PeriodicBeamCase
    kw = dict(n=8, dt=0.002, tf=0.02, orderphi=3, ordereta=3, k=2)
    out = run_periodic_beam(Periodic_Beam_params(**kw), device=0)
    r = PeriodicBeamCase(**kw).run()
    E = np.array(r["E"])
    assert len(out[10]) == len(r["t"]) and np.allclose(out[10], r["t"], rtol=1e-14)
    assert np.abs(out[0] - np.array(r["e_phi"])).max() <= 1e-6 * np.max(r["e_phi"])
    assert np.abs(out[1] - np.array(r["e_eta"])).max() <= 1e-6 * np.max(r["e_eta"])
    for k in range(4):
        assert np.abs(out[2 + k] - E[:, k]).max() <= 1e-8 * np.abs(E[:, k]).max(), k
    assert out[6] == r["E0"]["E_kin_f0"] and out[7] == r["E0"]["E_kin_s0"]


@pytest.mark.parametrize("omega", [0.4, 0.8])                                 # scripts/5-3-1-Liu.jl:20-32
def test_liu(omega, golden_dir):
    """5-3-1 on the reference's own model file (6 016 P4 triangles, variable bathymetry): matrix and
    right-hand side against the oracle, solution against its LU, |η|/η0 against Liu et al.'s curve."""
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.liu import LiuProblem, Liu_params, run_Liu
    from oracle.cases.liu import LiuCase
    f = os.path.join(golden_dir, "Liu.json")
    prob = LiuProblem(Liu_params(omega=omega, mesh_file=f), device=0)
    prob.assemble()
    ref = LiuCase(f, omega=omega)
    A, b = ref.assemble()
    assert prob.sys.ncoo == ref.ncoo
    A = _check_matrix(prob.sys, 0, A)
    assert np.abs(prob.sys.get_vector(0) - b).max() <= 1e-12 * np.abs(b).max()
    prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    x, st = prob.sys.solve(vec=0)
    xr = lu_reference(A, b)
    assert np.linalg.norm(x - xr) <= 1e-8 * np.linalg.norm(xr)
    xs, eta = run_Liu(Liu_params(omega=omega, mesh_file=f), device=0)
    xs_r, eta_r = ref.postprocess(xr)
    assert np.array_equal(xs, xs_r) and np.abs(eta - eta_r).max() <= 1e-7 * eta_r.max()
    lit = np.loadtxt(os.path.join(golden_dir, "Liu_omega_0%d.csv" % round(omega * 10)), delimiter=",")
    xu, iu = np.unique(xs, return_index=True)
    assert np.abs(np.interp(lit[:, 0], xu, eta[iu]) - lit[:, 1]).max() < 0.02


@pytest.mark.parametrize("kw", [dict(n=4, dt=0.01, tf=0.05, order=2, k=1),    # scripts/5-1-4-...jl:37-46 (coarse)
                                dict(n=6, dt=1e-3, tf=1e-2, order=4, k=3)])
def test_periodic_beam_fs(kw):
    """5-1-4: K, C, M and the Newmark matrix against the oracle, the stepped states, and the
    reference's tuple of errors and energies with the GPU functionals (src/Periodic_Beam_FS.jl:191-198)."""
    from vlfs_b200.cases.periodic_beam_fs import PeriodicBeamFSProblem, Periodic_Beam_FS_params, run_periodic_beam_FS
    from vlfs_b200.cases.periodic_beam import MAT_K, MAT_C, MAT_M, MAT_A
    from oracle.cases.periodic_beam_fs import PeriodicBeamFSCase
    prob = PeriodicBeamFSProblem(Periodic_Beam_FS_params(**kw), device=0)
    prob.assemble()
    ref = PeriodicBeamFSCase(**kw)
    K, C, M = ref.assemble()
    for mat, A in ((MAT_K, K), (MAT_C, C), (MAT_M, M)):
        _check_matrix(prob.sys, mat, A)
    cu, cuu = 0.5 / (0.25 * kw["dt"]), 1.0 / (0.25 * kw["dt"] ** 2)
    _check_matrix(prob.sys, MAT_A, K + cu * C + cuu * M)
    u, v, a, states, st = prob.solve()
    r = ref.run(keep_states=True)
    sref = np.array(r["states"])
    assert states.shape == sref.shape
    assert np.abs(states - sref).max() <= 1e-8 * np.abs(sref).max()
    out = run_periodic_beam_FS(Periodic_Beam_FS_params(**kw), device=0)
    E = np.array(r["E"])
    assert np.allclose(out[10], r["t"], rtol=1e-14)
    assert np.abs(out[0] - np.array(r["e_phi"])).max() <= 1e-6 * np.max(r["e_phi"])
    assert np.abs(out[1] - np.array(r["e_eta"])).max() <= 1e-6 * np.max(r["e_eta"])
    for k in range(4):
        assert np.abs(out[2 + k] - E[:, k]).max() <= 1e-7 * np.abs(E[:, k]).max(), k
    assert out[6] == r["E0"]["E_kin_f0"] and out[7] == r["E0"]["E_kin_s0"]
    assert out[8] == r["E0"]["E_pot_f0"] and out[9] == r["E0"]["E_ela_s0"]
