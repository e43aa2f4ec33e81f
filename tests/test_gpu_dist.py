"""Row-partitioned path on the GPU: 2-4 ranks run in lockstep inside one process on one B200
(one libvlfs_b200 context each; no rank ever waits on another's kernel), checked against the
oracle's global matrix and LU solution.  Real multi-GPU runs use the same `Fleet` code with
torch.distributed (scripts/dist_check.py, bench.py --gpus N)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla
from conftest import lu_reference

pytestmark = pytest.mark.gpu

KW = dict(nx=2, ny=1, nz=2, order=4, lfactor=0.6, dfactor=4)


def _fleet(nranks, block_lu=True):
    from vlfs_b200 import dist as D, _lib as L
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    members, probs = [], []
    for r in range(nranks):
        mk = lambda *a, **k: D.DistributedSystem(*a, rank=r, nranks=nranks, **k)
        prob = YagoProblem(Yago_freq_domain_params(**KW), system_cls=mk)
        ops = D.GpuOps(prob.sys)
        ops.bind_stream()
        prob.assemble()
        if block_lu:
            prob.sys.local.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=2, maxit=2, rtol=1e-12)
        members.append(D.Member(prob.sys, ops))
        probs.append(prob)
    return D.Fleet(members, nranks), members, probs


@pytest.fixture(scope="module")
def reference():
    from oracle.cases.yago import YagoCase
    ref = YagoCase(**KW)
    A, b = ref.assemble()
    return A.tocsr(), b


@pytest.mark.parametrize("nranks", [2, 3])
def test_partitioned_assembly_and_spmv(nranks, reference):
    A, b = reference
    N = A.shape[0]
    fleet, members, probs = _fleet(nranks)
    x = np.sin(0.1 * np.arange(N)) + 1j * np.cos(0.1 * np.arange(N))
    xs = []
    for m in members:
        xl = m.d.to_local(x)
        xl[m.d.n_owned:] = 0
        xs.append(m.ops.from_host(xl))
    ys = [m.ops.zeros() for m in members]
    fleet.matvec(xs, ys)
    y = np.zeros(N, dtype=complex)
    for m, yl in zip(members, ys):
        m.d.scatter_owned(m.ops.to_host(yl), y)
        Al = m.d.local.get_matrix(0)[:m.d.n_owned]
        G = A[m.d.owned][:, m.d.local_gids]
        assert abs(Al - G).max() <= 1e-12 * abs(A).max()
        bl = m.d.local.get_vector(0)[:m.d.n_owned]
        assert np.abs(bl - b[m.d.owned]).max() <= 1e-12 * np.abs(b).max()
    yr = A @ x
    assert np.abs(y - yr).max() <= 1e-12 * np.abs(yr).max()
    assert fleet.halo_bytes > 0


@pytest.mark.parametrize("nranks", [2, 4])
def test_distributed_gmres_block_lu(nranks, reference):
    A, b = reference
    fleet, members, probs = _fleet(nranks)
    bs = [m.ops.rhs_vector() for m in members]
    xs = [m.ops.zeros() for m in members]
    st = fleet.gmres(bs, xs, restart=80, maxit=400, rtol=1e-11)
    print(nranks, st)
    assert st["converged"], st
    x = np.zeros(len(b), dtype=complex)
    for m, xl in zip(members, xs):
        m.d.scatter_owned(m.ops.to_host(xl), x)
    xr = lu_reference(A, b)
    assert np.linalg.norm(x - xr) <= 1e-8 * np.linalg.norm(xr)


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_substructuring_exact_solve(nranks, reference):
    """Partial multifrontal factorisations with a kept interface + dense interface LU: one
    application is the LU-quality solution; as GMRES preconditioner it converges at once."""
    from vlfs_b200 import dist as D
    A, b = reference
    fleet, members, probs = _fleet(nranks, block_lu=False)
    schur = D.SchurSolver(fleet)
    schur.setup()
    bs = [m.ops.rhs_vector() for m in members]
    xs = [m.ops.zeros() for m in members]
    st = fleet.gmres(bs, xs, restart=6, maxit=6, rtol=1e-11, apply=schur.apply)
    print(nranks, "interface", schur.ng, st)
    assert st["converged"] and st["iterations"] <= 3, st
    x = np.zeros(len(b), dtype=complex)
    for m, xl in zip(members, xs):
        m.d.scatter_owned(m.ops.to_host(xl), x)
    xr = lu_reference(A, b)
    assert np.linalg.norm(x - xr) <= 1e-8 * np.linalg.norm(xr)
