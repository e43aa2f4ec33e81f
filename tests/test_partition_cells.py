"""Cell-based partition and halo plan of the multi-GPU path (host logic, CPU only): recursive
coordinate bisection of the cells, DOF ownership, the `vlfs_dist_set_halo` arrays every rank derives
without communication, and the chain structure the substructuring solver needs."""
import numpy as np
import pytest
from vlfs_b200 import dist as D
from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
from oracle.dist_cpu import CpuLocalSystem

KW = dict(nx=2, ny=1, nz=1, order=2, lfactor=0.4, dfactor=3)


def _rank(r, n, **kw):
    mk = lambda *a, **k: D.DistributedSystem(*a, rank=r, nranks=n, backend=CpuLocalSystem, partition="cells",
                                             rcb_axes=(0,), **{**k, **kw})
    return YagoProblem(Yago_freq_domain_params(**KW), system_cls=mk).sys


def test_rcb_tree_balances_and_classifies():
    rng = np.random.default_rng(0)
    P = rng.random((1000, 3)) * np.array([17.0, 5.0, 3.0])
    for n in (2, 3, 5, 8):
        part = D.rcb_classify(D.rcb_tree(P, n), P)
        cnt = np.bincount(part, minlength=n)
        assert cnt.min() >= 1000 // n - 2 and cnt.max() <= 1000 // n + 2 + n
    # restricted to x: slabs, part numbers ascend with x
    part = D.rcb_classify(D.rcb_tree(P, 4, axes=(0,)), P)
    order = np.argsort(P[:, 0])
    assert np.all(np.diff(part[order]) >= 0)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_halo_plans_agree_between_ranks_and_form_a_chain(n):
    ranks = [_rank(r, n) for r in range(n)]
    N = ranks[0].ndofs_global
    owned = np.concatenate([s.owned for s in ranks])
    assert np.array_equal(np.sort(owned), np.arange(N))
    plans = [D.local_halo_arrays(s) for s in ranks]
    for r, (s, (peers, sp, si, rp)) in enumerate(zip(ranks, plans)):
        assert rp[-1] == s.n_ghost and sp[0] == 0 and rp[0] == 0
        for k, q in enumerate(peers):
            mine_from_q = s.ghosts[rp[k]:rp[k + 1]]                   # what I receive from q, in my ghost order
            assert np.all(s.owner[mine_from_q] == q)
            pq, spq, siq, rpq = plans[q]
            kk = list(pq).index(r)
            sent = ranks[q].local_gids[siq[spq[kk]:spq[kk + 1]]]      # what q sends to me
            assert np.array_equal(sent, mine_from_q)
        # chain: cut unknowns (owned, seen by a lower rank) only go to rank-1; kept ghosts come from rank+1
        for k, q in enumerate(peers):
            if q < r and sp[k + 1] > sp[k]:
                assert q == r - 1
            if q > r and rp[k + 1] > rp[k]:
                assert q == r + 1


def test_cell_partition_gives_the_minimal_cut():
    """Cuts on cell faces: the unknowns rank 1 shares with rank 0 are one plane of ϕ DOFs and one line of
    κ DOFs - less than half of what equal-DOF-count slabs through the middle of cells need."""
    s_cells = _rank(1, 2)
    mk = lambda *a, **k: D.DistributedSystem(*a, rank=1, nranks=2, backend=CpuLocalSystem, partition="dofs", **k)
    s_dofs = YagoProblem(Yago_freq_domain_params(**KW), system_cls=mk).sys
    cut = lambda s: int((s.owner[np.unique(np.concatenate([t.ghosts for t in [_other(s)]]))] == 1).sum())

    def _other(s):
        part = "cells" if s is s_cells else "dofs"
        mk0 = lambda *a, **k: D.DistributedSystem(*a, rank=0, nranks=2, backend=CpuLocalSystem, partition=part,
                                                  rcb_axes=(0,), **k)
        return YagoProblem(Yago_freq_domain_params(**KW), system_cls=mk0).sys
    c_cells, c_dofs = cut(s_cells), cut(s_dofs)
    assert 0 < c_cells < c_dofs
    x = s_cells.coords[:, 0]
    g = _other(s_cells).ghosts
    assert len(np.unique(np.round(x[g[s_cells.owner[g] == 1]], 9))) == 1          # one x-plane
