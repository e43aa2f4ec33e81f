"""Committed oracle fixtures (tests/golden/oracle_*.npz, made by tests/golden/make_golden.py):
the CPU half checks that the oracle still reproduces them; the GPU half checks the CUDA path
against them without re-deriving anything from the oracle at run time."""
import os
import numpy as np
import pytest
from golden.make_golden import pattern_digest, probe


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _check_probe(g, A, b, tol=1e-12):
    p = probe(A, b)
    assert p["n"] == int(g["n"]) and p["nnz"] == int(g["nnz"])
    assert np.array_equal(p["digest"], g["digest"])                       # pattern bit-exact
    assert np.abs(p["Ax"] - g["Ax"]).max() <= tol * np.abs(g["Ax"]).max()
    assert np.abs(p["diag"] - g["diag"]).max() <= tol * np.abs(g["diag"]).max()
    assert np.abs(p["b"] - g["b"]).max() <= tol * np.abs(g["b"]).max()


def test_oracle_reproduces_fixtures(golden_dir):
    from oracle.cases.yago import YagoCase
    from oracle.cases.khabakhpasheva import KhabakhpashevaFreqCase
    from oracle.cases.multigeo import MultiGeoCase
    _check_probe(_load(golden_dir, "oracle_yago_warmup.npz"),
                 *YagoCase(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3).assemble(), tol=1e-13)
    _check_probe(_load(golden_dir, "oracle_khab_freq_warmup.npz"),
                 *KhabakhpashevaFreqCase(nx=10, ny=1, order=2, xi=0.0).assemble(), tol=1e-13)
    _check_probe(_load(golden_dir, "oracle_multigeo_warmup.npz"),
                 *MultiGeoCase(os.path.join(golden_dir, "multi_geo_coarse.json"), order=2).assemble(), tol=1e-13)


def _check_liu(g, A, b, tol):
    p = probe(A, b)
    assert p["n"] == int(g["n"]) and p["nnz"] == int(g["nnz"]) and np.array_equal(p["digest"], g["digest"])
    for key in ("Ax", "diag", "b"):
        assert np.abs(p[key][::16] - g[key]).max() <= tol * np.abs(g[key]).max(), key


def test_oracle_reproduces_liu_and_fs_fixtures(golden_dir):
    from oracle.cases.liu import LiuCase
    from oracle.cases.periodic_beam_fs import PeriodicBeamFSCase
    _check_liu(_load(golden_dir, "oracle_liu_omega04.npz"),
               *LiuCase(os.path.join(golden_dir, "Liu.json"), omega=0.4).assemble(), tol=1e-13)
    g = _load(golden_dir, "oracle_periodic_beam_fs.npz")
    r = PeriodicBeamFSCase(n=4, dt=0.01, tf=0.05, order=2, k=1).run()
    assert np.abs(r["u"] - g["u"]).max() <= 1e-12 * np.abs(g["u"]).max()
    assert np.abs(np.array(r["E"]) - g["E"]).max() <= 1e-11 * np.abs(g["E"]).max()


@pytest.mark.gpu
def test_gpu_liu_and_fs_against_fixtures(golden_dir):
    from vlfs_b200.cases.liu import LiuProblem, Liu_params, run_Liu
    from vlfs_b200.cases.periodic_beam_fs import run_periodic_beam_FS, Periodic_Beam_FS_params
    g = _load(golden_dir, "oracle_liu_omega04.npz")
    par = Liu_params(omega=0.4, mesh_file=os.path.join(golden_dir, "Liu.json"))
    prob = LiuProblem(par)
    prob.assemble()
    _check_liu(g, prob.sys.get_matrix(0), prob.sys.get_vector(0), tol=1e-12)
    xs, eta = run_Liu(par)
    assert np.array_equal(xs, g["xs"]) and np.abs(eta - g["eta"]).max() <= 1e-7 * g["eta"].max()
    g = _load(golden_dir, "oracle_periodic_beam_fs.npz")
    out = run_periodic_beam_FS(Periodic_Beam_FS_params(n=4, dt=0.01, tf=0.05, order=2, k=1))
    assert np.abs(out[0] - g["e_phi"]).max() <= 1e-7 * g["e_phi"].max()
    assert np.abs(out[1] - g["e_eta"]).max() <= 1e-7 * g["e_eta"].max()
    for k in range(4):
        assert np.abs(out[2 + k] - g["E"][:, k]).max() <= 1e-7 * np.abs(g["E"][:, k]).max(), k


@pytest.mark.gpu
def test_gpu_yago_warmup_against_fixture(golden_dir):
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    g = _load(golden_dir, "oracle_yago_warmup.npz")
    prob = YagoProblem(Yago_freq_domain_params(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3), device=0)
    prob.assemble()
    _check_probe(g, prob.sys.get_matrix(0), prob.sys.get_vector(0))
    prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    x, st = prob.sys.solve(vec=0)
    assert np.linalg.norm(x - g["x"]) <= 1e-8 * np.linalg.norm(g["x"])
    xs, rao = prob.rao(x)
    assert np.array_equal(xs, g["rao_x"]) and np.abs(rao - g["rao"]).max() <= 1e-8 * np.abs(g["rao"]).max()


@pytest.mark.gpu
def test_gpu_khabakhpasheva_against_fixtures(golden_dir):
    from vlfs_b200.cases.khabakhpasheva import (run_Khabakhpasheva_freq_domain, run_Khabakhpasheva_time_domain,
                                                Khabakhpasheva_freq_domain_params, Khabakhpasheva_time_domain_params,
                                                KhabakhpashevaTimeProblem)
    g = _load(golden_dir, "oracle_khab_freq_warmup.npz")
    xs, eta = run_Khabakhpasheva_freq_domain(Khabakhpasheva_freq_domain_params(nx=10, ny=1, order=2, xi=0.0))
    assert np.array_equal(xs, g["xs"]) and np.abs(eta - g["eta"]).max() <= 1e-8 * g["eta"].max()
    g = _load(golden_dir, "oracle_khab_time_warmup.npz")
    prob = KhabakhpashevaTimeProblem(Khabakhpasheva_time_domain_params(nx=10, ny=1, order=2, xi=0.0))
    prob.assemble()
    ts, hist, (u, v, a), st = prob.solve(nsteps=40)
    assert np.allclose(ts, g["ts"], rtol=1e-14, atol=0)
    assert np.abs(hist - g["eta_hist"]).max() <= 1e-8 * np.abs(g["eta_hist"]).max()
    assert np.abs(u - g["u"]).max() <= 1e-8 * np.abs(g["u"]).max()


@pytest.mark.gpu
def test_gpu_periodic_beam_and_multigeo_against_fixtures(golden_dir):
    from vlfs_b200.cases.periodic_beam import run_periodic_beam, Periodic_Beam_params
    from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
    from vlfs_b200 import _lib as L
    g = _load(golden_dir, "oracle_periodic_beam.npz")
    out = run_periodic_beam(Periodic_Beam_params(n=3, dt=0.5, tf=1.0, orderphi=2, ordereta=2, k=1))
    assert np.abs(out[0] - g["e_phi"]).max() <= 1e-7 * g["e_phi"].max()
    assert np.abs(out[1] - g["e_eta"]).max() <= 1e-7 * g["e_eta"].max()
    out = run_periodic_beam(Periodic_Beam_params(n=8, dt=1e-6, tf=1e-4, orderphi=2, ordereta=2, k=15))
    assert abs(out[0][-1] - g["conv_e_phi"]) <= 1e-6 * g["conv_e_phi"]
    assert abs(out[1][-1] - g["conv_e_eta"]) <= 1e-6 * g["conv_e_eta"]
    g = _load(golden_dir, "oracle_multigeo_warmup.npz")
    prob = MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=os.path.join(golden_dir, "multi_geo_coarse.json"), order=2))
    prob.assemble()
    _check_probe(g, prob.sys.get_matrix(0), prob.sys.get_vector(0), tol=2e-12)
    prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    x, st = prob.sys.solve(vec=0)
    assert np.linalg.norm(x - g["x"]) <= 1e-8 * np.linalg.norm(g["x"])
