import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def lu_reference(A, b, steps=3):
    """The oracle's `LUSolver()` solution (SuperLU) followed by iterative refinement with the
    residual accumulated in extended precision.  The monolithic matrices reach condition numbers
    of 1e11 (MultiGeo order 4: 2.6e11), where a plain double-precision LU solution is itself only
    accurate to ~5e-9; the 1e-8 parity gate is applied against the refined solution."""
    import numpy as np
    import scipy.sparse.linalg as spla
    A = A.tocsc()
    lu = spla.splu(A)
    x = lu.solve(b)
    coo = A.tocoo()
    ext = np.clongdouble if np.iscomplexobj(x) or np.iscomplexobj(A.data) else np.longdouble
    xe = x.astype(ext)
    for _ in range(steps):
        r = b.astype(ext).copy()
        np.subtract.at(r, coo.row, coo.data.astype(ext) * xe[coo.col])
        xe = xe + lu.solve(r.astype(x.dtype)).astype(ext)
    return xe.astype(x.dtype)
