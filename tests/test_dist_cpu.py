"""Multi-rank host logic (partition, local numbering, halo plan, distributed GMRES with
block-Jacobi/LU) on the CPU: in-process fleets of 2-4 ranks and a world_size-2 gloo run.  The
arithmetic twin is oracle/dist_cpu.py; the GPU path runs the same `Fleet` code on `GpuOps`."""
import os
import sys
import numpy as np
import pytest
import scipy.sparse.linalg as spla
import vlfs_b200
from vlfs_b200 import dist as D
from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
from oracle.dist_cpu import CpuLocalSystem, CpuOps
from oracle.cases.yago import YagoCase

KW = dict(nx=2, ny=1, nz=1, order=2, lfactor=0.4, dfactor=3)


def _member(rank, nranks):
    mk = lambda *a, **k: D.DistributedSystem(*a, rank=rank, nranks=nranks, backend=CpuLocalSystem, **k)
    prob = YagoProblem(Yago_freq_domain_params(**KW), system_cls=mk)
    prob.assemble()
    return prob, D.Member(prob.sys, CpuOps(prob.sys))


def test_slab_owner_balanced_and_tie_safe():
    x = np.repeat(np.arange(10.0), 7)
    o = D.slab_owner(x, 4)
    assert o.min() == 0 and o.max() == 3 and np.all(np.diff(o) >= 0)
    for v in range(10):
        assert len(set(o[x == v])) == 1
    assert np.array_equal(D.slab_owner(x, 1), np.zeros(70))


@pytest.mark.parametrize("nranks", [2, 3])
def test_local_systems_reproduce_global_rows_and_spmv(nranks):
    ref = YagoCase(**KW)
    A, b = ref.assemble()
    A = A.tocsr()
    N = A.shape[0]
    x = np.sin(0.1 * np.arange(N)) + 1j * np.cos(0.1 * np.arange(N))
    members = [_member(r, nranks)[1] for r in range(nranks)]
    owned_all = np.concatenate([m.d.owned for m in members])
    assert np.array_equal(np.sort(owned_all), np.arange(N))              # a partition
    fleet = D.Fleet(members, nranks)
    xs = []
    for m in members:
        v = m.ops.from_host(m.d.to_local(x))
        v[m.d.n_owned:] = 0                                              # ghosts come from the halo
        xs.append(v)
    ys = [m.ops.zeros() for m in members]
    fleet.matvec(xs, ys)
    y = np.zeros(N, dtype=complex)
    for m, yl in zip(members, ys):
        m.d.scatter_owned(yl, y)
        # owned rows of the local matrix == rows of the global matrix
        Al = m.ops.A.tocsr()
        G = A[m.d.owned][:, m.d.local_gids]
        assert abs(Al - G).max() <= 1e-12 * abs(A).max()
        bl = m.ops.rhs_vector()[:m.d.n_owned]
        assert np.abs(bl - b[m.d.owned]).max() <= 1e-12 * np.abs(b).max()
    assert np.abs(y - A @ x).max() <= 1e-12 * np.abs(A @ x).max()


@pytest.mark.parametrize("nranks", [2, 4])
def test_distributed_gmres_matches_lu(nranks):
    ref = YagoCase(**KW)
    A, b = ref.assemble()
    xr = spla.splu(A.tocsc()).solve(b)
    members = [_member(r, nranks)[1] for r in range(nranks)]
    fleet = D.Fleet(members, nranks)
    bs = [m.ops.rhs_vector() for m in members]
    xs = [m.ops.zeros() for m in members]
    st = fleet.gmres(bs, xs, restart=60, maxit=300, rtol=1e-11)
    assert st["converged"], st
    x = np.zeros(len(b), dtype=complex)
    for m, xl in zip(members, xs):
        m.d.scatter_owned(xl, x)
    assert np.linalg.norm(x - xr) <= 1e-8 * np.linalg.norm(xr)


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_substructuring_is_an_exact_solve(nranks):
    """SchurSolver: partial factorisations + interface system = A^-1 (one application)."""
    ref = YagoCase(**KW)
    A, b = ref.assemble()
    xr = spla.splu(A.tocsc()).solve(b)
    members = [_member(r, nranks)[1] for r in range(nranks)]
    fleet = D.Fleet(members, nranks)
    schur = D.SchurSolver(fleet)
    schur.setup()
    assert 0 < schur.ng < 0.2 * len(b)
    bs = [m.ops.rhs_vector() for m in members]
    zs = [m.ops.zeros() for m in members]
    schur.apply(bs, zs)
    x = np.zeros(len(b), dtype=complex)
    for m, zl in zip(members, zs):
        m.d.scatter_owned(zl, x)
    assert np.linalg.norm(x - xr) <= 1e-9 * np.linalg.norm(xr)
    # as the preconditioner of the distributed GMRES: converges at once
    xs = [m.ops.zeros() for m in members]
    st = fleet.gmres(bs, xs, restart=5, maxit=5, rtol=1e-11, apply=schur.apply)
    assert st["converged"] and st["iterations"] <= 2, st


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob, member = _member(rank, world)
        fleet = D.Fleet([member], world, dist=dist)
        bs = [member.ops.rhs_vector()]
        xs = [member.ops.zeros()]
        # torch.distributed needs tensors: wrap the NumPy buffers (shared memory, CPU)
        st = fleet.gmres(bs, xs, restart=60, maxit=300, rtol=1e-11)
        schur = D.SchurSolver(fleet)                     # exact substructuring over gloo
        schur.setup()
        xs2 = [member.ops.zeros()]
        st2 = fleet.gmres(bs, xs2, restart=5, maxit=5, rtol=1e-11, apply=schur.apply)
        assert st2["converged"] and st2["iterations"] <= 2, st2
        assert np.abs(xs2[0][:member.d.n_owned] - xs[0][:member.d.n_owned]).max() <= 1e-8 * np.abs(xs[0]).max()
        q.put((rank, member.d.owned, np.asarray(xs2[0][:member.d.n_owned]), st))
    except Exception as e:          # surface the failure instead of a queue timeout
        q.put((rank, None, repr(e), None))
        raise
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    assert all(o[1] is not None for o in out), out
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = YagoCase(**KW)
    A, b = ref.assemble()
    xr = spla.splu(A.tocsc()).solve(b)
    x = np.zeros(len(b), dtype=complex)
    for rank, owned, xl, st in out:
        assert st["converged"]
        x[owned] = xl
    assert np.linalg.norm(x - xr) <= 1e-8 * np.linalg.norm(xr)
