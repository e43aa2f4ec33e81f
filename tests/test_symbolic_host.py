"""Host side of the sparse LU analysis (csrc/symbolic_host.hpp: A + A^T, geometric nested dissection,
update sets) compiled with plain g++ and run on oracle matrices: the elimination tree must be valid
(permutation bijective, children before parents, every later neighbour of a pivot inside the front,
update sets nested in the parent's front) and independent of the number of host threads - CPU only."""
import json
import os
import struct
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("sym") / "symbolic_bench")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "monolithicfemvlfs.jl_b200", "csrc"),
                    os.path.join(ROOT, "scripts", "symbolic_bench.cpp"), "-o", exe], check=True)
    return exe


def _write(path, A, xy, role, leaf):
    A = A.tocsr(); A.sort_indices()
    n = A.shape[0]
    with open(path, "wb") as f:
        f.write(struct.pack("<qqii", n, A.nnz, xy.shape[1], leaf))
        f.write(A.indptr.astype(np.int32).tobytes())
        f.write(A.indices.astype(np.int32).tobytes())
        f.write(np.ascontiguousarray(xy, dtype=np.float64).tobytes())
        f.write(np.asarray(role, dtype=np.int8).tobytes())


def _run(exe, path, threads):
    env = dict(os.environ, VLFS_HOST_THREADS=str(threads))
    out = subprocess.run([exe, path], env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_analysis_is_valid_and_thread_independent(bench, tmp_path):
    from oracle.cases.yago import YagoCase
    c = YagoCase(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3)
    A, _ = c.assemble()
    xy = np.concatenate([s.dof_coords() for s in c.X.spaces])
    n = A.shape[0]
    f = str(tmp_path / "y0.bin")
    _write(f, A, xy, np.zeros(n), leaf=48)
    r1, r4 = _run(bench, f, 1), _run(bench, f, 4)
    assert r1["valid"] and r4["valid"] and r1["n"] == n and r1["n_keep"] == 0
    for key in ("fronts", "levels", "max_front", "fill", "pos_digest"):
        assert r1[key] == r4[key], key
    assert r1["fronts"] > 100 and r1["max_front"] < n // 4


def test_analysis_with_schur_boundary_and_ghosts(bench, tmp_path):
    """Roles of the substructuring solver: one x-layer kept as Schur boundary, the far end ignored."""
    from oracle.cases.khabakhpasheva import KhabakhpashevaFreqCase
    c = KhabakhpashevaFreqCase(nx=10, ny=2, order=2, xi=0.0)
    A, _ = c.assemble()
    xy = np.concatenate([s.dof_coords() for s in c.X.spaces])
    x = xy[:, 0]
    role = np.zeros(A.shape[0], dtype=np.int8)
    role[x > 0.8 * x.max()] = 2
    role[(x > 0.5 * x.max()) & (x <= 0.52 * x.max())] = 1
    f = str(tmp_path / "k.bin")
    _write(f, A, xy, role, leaf=32)
    r = _run(bench, f, 3)
    assert r["valid"] and r["n_keep"] == int((role == 1).sum()) and r["n"] == int((role == 0).sum())
