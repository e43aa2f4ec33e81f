"""The multi-GPU path behind the C ABI (csrc/dist.cu): cell partition, halo exchange overlapped with
the interior rows, one all-reduce per Krylov iteration, substructuring on the chain of slab cuts.
On a box with one GPU the ranks share the device and the library uses its in-process transport
(same kernels, same chain algebra, same GMRES); with >= 2 GPUs the same test runs over NCCL."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu

KW = dict(nx=2, ny=1, nz=1, order=2, lfactor=0.4, dfactor=3)


def _single():
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    p = YagoProblem(Yago_freq_domain_params(**KW), device=0)
    p.assemble()
    p.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=10, maxit=20, rtol=1e-12)
    x, st = p.sys.solve(vec=0)
    return p, x


def _multi(nranks, devices, kw=KW):
    from vlfs_b200 import dist as D
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    mk = lambda *a, **k: D.MultiSystem(*a, nranks=nranks, devices=devices, **{kk: v for kk, v in k.items() if kk != "device"})
    p = YagoProblem(Yago_freq_domain_params(**kw), system_cls=mk)
    p.assemble()
    return p


def _check_against_single(pm, nranks):
    ps, xs = _single()
    A = ps.sys.get_matrix(0)
    b = ps.sys.get_vector(0)
    N = A.shape[0]
    # right-hand side and product: owner-computes rows are complete, the halo carries the rest
    bm = pm.sys.get_vector(0)
    assert np.abs(bm - b).max() <= 1e-12 * np.abs(b).max()
    x = np.sin(0.1 * np.arange(N)) + 1j * np.cos(0.1 * np.arange(N))
    ym = pm.sys.spmv(x)
    yr = A @ x
    assert np.abs(ym - yr).max() <= 1e-12 * np.abs(yr).max()
    # exact distributed solve
    pm.sys.solver_setup(0, restart=10, maxit=20, rtol=1e-12)
    xm, st = pm.sys.solve(vec=0)
    assert st["converged"], st
    assert st["iterations"] <= 4                                   # exact preconditioner: refinement only
    assert np.linalg.norm(xm - xs) / np.linalg.norm(xs) < 1e-9
    assert np.linalg.norm(A @ xm - b) / np.linalg.norm(b) < 1e-11
    info = pm.sys.dist_info()
    assert info[0]["n_left"] == 0 and info[-1]["n_right"] == 0
    for r in range(nranks - 1):
        assert info[r]["n_right"] == info[r + 1]["n_left"] > 0      # neighbouring links share a cut
    # the ω sweep: new coefficients, refactorise, solve again
    pm.set_frequency(0.6); pm.assemble(); pm.sys.refactor(0)
    ps.set_frequency(0.6); ps.assemble(); ps.sys.refactor(0)
    xm2, st2 = pm.sys.solve(vec=0)
    xs2, _ = ps.sys.solve(vec=0)
    assert st2["converged"] and np.linalg.norm(xm2 - xs2) / np.linalg.norm(xs2) < 1e-9
    kap_m, kap_s = pm.X.split(xm2)[1], ps.X.split(xs2)[1]        # free-surface elevation field
    assert len(kap_m) > 0 and np.abs(np.abs(kap_m) - np.abs(kap_s)).max() < 1e-9 * np.abs(kap_s).max()


@pytest.mark.parametrize("nranks", [2, 3])
def test_ranks_sharing_one_gpu_reproduce_the_single_gpu_solve(nranks):
    pm = _multi(nranks, [0] * nranks)
    _check_against_single(pm, nranks)


def test_two_gpus_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    pm = _multi(2, [0, 1])
    _check_against_single(pm, 2)


def test_four_ranks_on_a_longer_mesh_meet_the_solution_gate():
    """nx = 4: four slabs, three cuts; solution against SuperLU + extended-precision refinement."""
    from conftest import lu_reference
    kw = dict(KW, nx=4)
    import torch
    devs = [0, 1, 2, 3] if torch.cuda.device_count() >= 4 else [0] * 4
    pm = _multi(4, devs, kw)
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    ps = YagoProblem(Yago_freq_domain_params(**kw), device=0)
    ps.assemble()
    A, b = ps.sys.get_matrix(0), ps.sys.get_vector(0)
    xr = lu_reference(A, b)
    pm.sys.solver_setup(0, restart=10, maxit=20, rtol=1e-11)
    xm, st = pm.sys.solve(vec=0)
    assert st["converged"], st
    assert np.linalg.norm(xm - xr) / np.linalg.norm(xr) < 1e-8
