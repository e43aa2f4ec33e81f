"""Zero-fill incomplete LU, sequential restatement (test infrastructure - the checker of `ilu0.cu`).

north_star (5) names "Jacobi or block-ILU0" as the Krylov preconditioners; the reference itself only
ever calls `LUSolver()` (src/Periodic_Beam.jl:107), so this is the textbook IKJ algorithm (Saad,
*Iterative Methods for Sparse Linear Systems*, Alg. 10.4) on the CSR pattern of the leading
`n_owned x n_owned` block:

    for i in rows:  for k < i in row i (ascending):  a_ik /= a_kk
                        for j > k in row k that row i also holds:  a_ij -= a_ik a_kj

`levels` restates the dependency levels the GPU schedule uses (row i after every row of its strict
lower pattern); the factors do not depend on the schedule, so parity is to rounding.  Only `tests/`
imports this module.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def ilu0(A, n_owned=None):
    """Returns (L, U) as CSR matrices with unit-diagonal L, of the leading block of the sorted CSR matrix A."""
    A = sp.csr_matrix(A)
    n = A.shape[0] if n_owned is None else int(n_owned)
    A = A[:n, :n].tocsr()
    A.sort_indices()
    rp, ci = A.indptr, A.indices
    lu = A.data.copy()
    diag = np.empty(n, dtype=np.int64)
    for i in range(n):
        lo, hi = rp[i], rp[i + 1]
        d = lo + np.searchsorted(ci[lo:hi], i)
        assert d < hi and ci[d] == i, "structurally zero diagonal"
        diag[i] = d
        for p in range(lo, d):
            k = ci[p]
            lik = lu[p] / lu[diag[k]]
            q0, q1 = diag[k] + 1, rp[k + 1]
            if q1 > q0:
                cols = ci[q0:q1]
                pos = p + 1 + np.searchsorted(ci[p + 1:hi], cols)
                ok = pos < hi
                ok[ok] = ci[pos[ok]] == cols[ok]
                lu[pos[ok]] -= lik * lu[q0:q1][ok]
            lu[p] = lik
    rows = np.repeat(np.arange(n), np.diff(rp))
    low = ci < rows
    Lm = sp.csr_matrix((lu[low], (rows[low], ci[low])), shape=(n, n)) + sp.identity(n, format="csr", dtype=lu.dtype)
    Um = sp.csr_matrix((lu[~low], (rows[~low], ci[~low])), shape=(n, n))
    return Lm.tocsr(), Um.tocsr()


def apply(Lm, Um, r):
    """z = U^-1 L^-1 r."""
    y = spla.spsolve_triangular(Lm, r, lower=True, unit_diagonal=True)
    return spla.spsolve_triangular(Um, y, lower=False)


def levels(A, n_owned=None):
    """(lower levels, upper levels) per row of the leading block."""
    A = sp.csr_matrix(A)
    n = A.shape[0] if n_owned is None else int(n_owned)
    A = A[:n, :n].tocsr()
    A.sort_indices()
    rp, ci = A.indptr, A.indices
    ll = np.zeros(n, dtype=np.int64)
    ul = np.zeros(n, dtype=np.int64)
    for i in range(n):
        c = ci[rp[i]:rp[i + 1]]
        c = c[c < i]
        if len(c):
            ll[i] = ll[c].max() + 1
    for i in range(n - 1, -1, -1):
        c = ci[rp[i]:rp[i + 1]]
        c = c[c > i]
        if len(c):
            ul[i] = ul[c].max() + 1
    return ll, ul
