"""Host analysis of the GPU ILU(0) (csrc/ilu0_host.hpp: diagonal positions, dependency levels, rows ordered by
level, launch plan, diagonal blocks) compiled with plain g++ and checked against the restatement `oracle/ilu0.py`
and against the invariants the kernels rely on - CPU only."""
import json
import os
import struct
import subprocess
import numpy as np
import pytest
import scipy.sparse as sp
from oracle import ilu0 as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("ilu") / "ilu0_host_harness")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "monolithicfemvlfs.jl_b200", "csrc"),
                    os.path.join(ROOT, "tests", "ilu0_host_harness.cpp"), "-o", exe], check=True)
    return exe


def _run(exe, tmp_path, A, block_rows, n_owned=None):
    A = A.tocsr(); A.sort_indices()
    n = A.shape[0] if n_owned is None else n_owned
    path = str(tmp_path / "a.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<qq", n, int(A.indptr[n])))
        f.write(A.indptr[:n + 1].astype(np.int32).tobytes())
        f.write(A.indices[:A.indptr[n]].astype(np.int32).tobytes())
    out = subprocess.run([exe, path, str(block_rows)], capture_output=True, text=True)
    return out.returncode, json.loads(out.stdout.strip().splitlines()[-1])


def _grid(n):
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    return (sp.kron(sp.eye(n), T) + sp.kron(T, sp.eye(n))).tocsr()


def _fe_pattern():
    from oracle.cases.yago import YagoCase
    A, _ = YagoCase(nx=1, ny=1, nz=1, order=2, lfactor=0.4, dfactor=3).assemble()
    return A.tocsr()


def _block_diagonal(A, B):
    A = A.tocsr()
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    keep = (rows // B) == (A.indices // B)
    return sp.csr_matrix((np.ones(keep.sum()), (rows[keep], A.indices[keep])), shape=A.shape)


def _check_schedule(S, level, n, block_rows):
    rows, lptr = np.array(S["rows"]), np.array(S["lptr"])
    assert sorted(rows.tolist()) == list(range(n))                       # a permutation of the rows
    if block_rows == 0:
        assert lptr[0] == 0 and lptr[-1] == n and S["nlevels"] == level.max() + 1
        for l in range(S["nlevels"]):
            seg = rows[lptr[l]:lptr[l + 1]]
            assert np.all(level[seg] == l) and np.all(np.diff(seg) > 0)   # one level, ascending row ids
        # the plan covers the levels in order; wide levels alone, runs of narrow levels (<= 32 rows) together
        nxt = 0
        for l0, l1, chain in S["plan"]:
            assert l0 == nxt and l1 > l0
            widths = np.diff(lptr[l0:l1 + 1])
            assert (chain == 1 and widths.max() <= 32) or (chain == 0 and l1 == l0 + 1 and widths[0] > 32)
            nxt = l1
        assert nxt == S["nlevels"]
        for a, b in zip(S["plan"][:-1], S["plan"][1:]):
            assert not (a[2] == 1 and b[2] == 1)                          # maximal runs
    else:
        blk = np.array(S["blk"])
        nb = -(-n // block_rows)
        assert len(blk) == nb + 1 and blk[-1] == len(lptr)
        most = 0
        for b in range(nb):
            r0, r1 = b * block_rows, min(n, (b + 1) * block_rows)
            lp = lptr[blk[b]:blk[b + 1]]
            assert lp[0] == r0 and lp[-1] == r1                           # the block's rows, terminated
            for l in range(len(lp) - 1):
                seg = rows[lp[l]:lp[l + 1]]
                assert len(seg) and np.all((seg >= r0) & (seg < r1)) and np.all(level[seg] == l)
            most = max(most, len(lp) - 1)
        assert S["nlevels"] == most


@pytest.mark.parametrize("name", ["grid", "fe", "tridiagonal"])
@pytest.mark.parametrize("block_rows", [0, 64, 100])
def test_levels_match_the_restatement_and_the_schedule_is_consistent(harness, tmp_path, name, block_rows):
    A = {"grid": lambda: _grid(23), "fe": _fe_pattern,
         "tridiagonal": lambda: sp.diags([np.ones(299), 2 * np.ones(300), np.ones(299)], [-1, 0, 1]).tocsr()}[name]()
    n = A.shape[0]
    rc, out = _run(harness, tmp_path, A, block_rows)
    assert rc == 0 and out["n"] == n and out["block_rows"] == block_rows
    Ab = A if block_rows == 0 else _block_diagonal(A, block_rows)
    ll, ul = O.levels(Ab)
    assert np.array_equal(out["level_lower"], ll) and np.array_equal(out["level_upper"], ul)
    A = A.tocsr(); A.sort_indices()
    assert all(A.indices[d] == i for i, d in enumerate(out["diag"]))
    _check_schedule(out["lower"], ll, n, block_rows)
    _check_schedule(out["upper"], ul, n, block_rows)


def test_owned_block_drops_ghost_columns(harness, tmp_path):
    A = _grid(20)
    rc, out = _run(harness, tmp_path, A, 0, n_owned=150)
    ll, ul = O.levels(A, 150)
    assert rc == 0 and np.array_equal(out["level_lower"], ll) and np.array_equal(out["level_upper"], ul)


def test_unusable_patterns_are_refused(harness, tmp_path):
    A = sp.csr_matrix(np.array([[1.0, 1.0, 0], [1.0, 0.0, 1.0], [0, 1.0, 1.0]]))      # a_11 structurally zero
    A.eliminate_zeros()
    rc, out = _run(harness, tmp_path, A, 0)
    assert rc == 1 and "zero diagonal in row 1" in out["error"]
    B = sp.csr_matrix((np.ones(4), np.array([1, 0, 0, 1]), np.array([0, 2, 4])), shape=(2, 2))   # unsorted row
    path = str(tmp_path / "b.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<qq", 2, 4) + B.indptr.astype(np.int32).tobytes() + B.indices.astype(np.int32).tobytes())
    out = subprocess.run([harness, path, "0"], capture_output=True, text=True)
    assert out.returncode == 1 and "ascending" in out.stdout
