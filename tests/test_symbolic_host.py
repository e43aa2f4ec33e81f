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


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_analysis_on_random_graphs_and_degenerate_coordinates(bench, tmp_path, seed):
    """Not meshes: random sparse (unsymmetric) patterns, random roles, coordinates with many ties or all
    equal, isolated unknowns - the elimination tree must still be valid."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    n = int(rng.integers(300, 3000))
    A = sp.random(n, n, density=float(rng.uniform(0.002, 0.02)), random_state=int(seed), format="csr")
    A = (A + sp.eye(n, format="csr")).tocsr()
    dim = int(rng.integers(1, 4))
    if seed == 0:
        xy = np.zeros((n, dim))                                   # every coordinate equal: one big leaf
    elif seed == 1:
        xy = rng.integers(0, 3, size=(n, dim)).astype(float)      # three planes per axis: massive ties
    else:
        xy = rng.random((n, dim))
    role = rng.choice([0, 0, 0, 0, 1, 2], size=n).astype(np.int8)
    role[:5] = 0
    f = str(tmp_path / "r.bin")
    _write(f, A, xy, role, leaf=int(rng.integers(8, 64)))
    r1, r3 = _run(bench, f, 1), _run(bench, f, 3)
    assert r1["valid"] and r3["valid"]
    assert r1["n"] == int((role == 0).sum()) and r1["n_keep"] == int((role == 1).sum())
    assert r1["pos_digest"] == r3["pos_digest"] and r1["fronts"] == r3["fronts"]
