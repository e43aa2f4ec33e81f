"""The ILU(0) restatement (`oracle/ilu0.py`, checker of the GPU preconditioner) pinned on the defining
properties of the factorisation: (LU)_ij = A_ij on the pattern, exact LU when there is no fill, and the
level schedule being a valid topological order."""
import numpy as np
import scipy.sparse as sp
from oracle import ilu0 as O


def _grid(n):
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    return (sp.kron(sp.eye(n), T) + sp.kron(T, sp.eye(n)) + 0.3j * sp.kron(sp.eye(n), T)).tocsr()


def test_product_of_the_factors_reproduces_the_matrix_on_its_pattern():
    A = _grid(12)
    Lm, Um = O.ilu0(A)
    R = (Lm @ Um - A).tocsr()
    P = A.copy(); P.data = np.ones(P.nnz)
    assert abs(R.multiply(P)).max() <= 1e-14 * abs(A).max()
    assert abs(R).max() > 1e-3                        # fill positions are dropped: incomplete, not exact
    assert sp.tril(Um, -1).nnz == 0 and sp.triu(Lm, 1).nnz == 0 and np.allclose(Lm.diagonal(), 1)


def test_no_fill_means_exact_lu_and_blocks_restrict():
    n = 50
    rng = np.random.default_rng(0)
    A = sp.diags([rng.standard_normal(n - 1), 3 + rng.random(n), rng.standard_normal(n - 1)], [-1, 0, 1]).tocsr()
    Lm, Um = O.ilu0(A)
    assert abs(Lm @ Um - A).max() <= 1e-14
    x = rng.standard_normal(n)
    assert np.abs(O.apply(Lm, Um, A @ x) - x).max() <= 1e-12
    Lb, Ub = O.ilu0(A, 20)
    assert abs(Lb @ Ub - A[:20, :20]).max() <= 1e-14


def test_levels_are_a_topological_order():
    A = _grid(9)
    ll, ul = O.levels(A)
    A = A.tocsr(); A.sort_indices()
    for i in range(A.shape[0]):
        c = A.indices[A.indptr[i]:A.indptr[i + 1]]
        assert all(ll[k] < ll[i] for k in c[c < i]) and all(ul[k] < ul[i] for k in c[c > i])
    assert ll.max() + 1 == 2 * 9 - 1                  # anti-diagonals of the grid
