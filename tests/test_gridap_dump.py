"""Pinning the oracle against Gridap itself ("parity unpinned" until table dumps from a
Gridap install exist).  `julia/dump_gridap_tables.jl` writes the dumps; when `tests/golden/gridap_dump/`
holds them, the oracle's DOF numbering and sparsity pattern must match bit for bit, entries to 1e-12
and the LU solution to 1e-8.  Without dumps only the file format is exercised (a round trip through
the same reader), so that the comparison code is known to work the day the dumps arrive."""
import os
import numpy as np
import pytest
import scipy.sparse.linalg as spla
from oracle import gridap_dump as GD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMPS = os.path.join(ROOT, "tests", "golden", "gridap_dump")


def _cases():
    from oracle.cases.yago import YagoCase
    from oracle.cases.khabakhpasheva import KhabakhpashevaFreqCase
    return {
        "yago_warmup": lambda: YagoCase(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3),
        "khabakhpasheva_freq_warmup": lambda: KhabakhpashevaFreqCase(nx=2, ny=1, order=2, xi=0.0),
    }


def test_dump_format_round_trip(tmp_path):
    case = _cases()["yago_warmup"]()
    A, b = case.assemble()
    A = A.tocsc()
    x = spla.splu(A).solve(b)
    xyz = np.concatenate([s.dof_coords() for s in case.X.spaces])
    GD.write_dump(str(tmp_path), "yago_warmup", A, b, x, case.X.spaces, xyz)
    dump = GD.read_dump(os.path.join(str(tmp_path), "yago_warmup"))
    out = GD.compare(case, dump)
    assert out["entry_err"] == 0.0 and out["nnz"] == A.nnz
    # the comparison notices a renumbering and a perturbed entry
    bad = dict(dump, fields=[dict(f, cell_dofs=f["cell_dofs"][::-1].copy()) for f in dump["fields"]])
    with pytest.raises(AssertionError, match="numbering"):
        GD.compare(case, bad)
    A2 = dump["A"].copy()
    A2.data[np.argmax(np.abs(A2.data))] *= 1 + 1e-9
    with pytest.raises(AssertionError, match="entries"):
        GD.compare(case, dict(dump, A=A2))


@pytest.mark.parametrize("name", ["yago_warmup", "khabakhpasheva_freq_warmup"])
def test_oracle_matches_gridap_dump(name):
    d = os.path.join(DUMPS, name)
    if not os.path.exists(os.path.join(d, "manifest.txt")):
        pytest.skip("no Gridap dump for %s (run julia/dump_gridap_tables.jl on a machine with Julia)" % name)
    out = GD.compare(_cases()[name](), GD.read_dump(d))
    assert out["entry_err"] <= 1e-12 and out["solution_err"] <= 1e-8
