"""Block Jacobi over R ranks with ILU(0) inside each rank's block (SURVEY.md §8e "Preconditioner": the iteration
count grows with the rank count - report it).  CPU study with the restatement `oracle/ilu0.py` and SciPy's GMRES on
the Newmark matrix of configuration 5-1-1 (rows cut into R consecutive blocks = what R GPUs owning consecutive rows
apply with VLFS_PRECOND_BLOCK_ILU0) and on an SPD grid operator.  `python scripts/ilu0_rank_study.py`"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from oracle import ilu0 as O


def block_diagonal(A, B):
    A = A.tocsr()
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    keep = (rows // B) == (A.indices // B)
    return sp.csr_matrix((A.data[keep], (rows[keep], A.indices[keep])), shape=A.shape)


def iterations(A, b, M, restart, maxit):
    it = [0]
    x, info = spla.gmres(A, b, M=M, rtol=1e-10, restart=restart, maxiter=maxit,
                         callback=lambda r: it.__setitem__(0, it[0] + 1), callback_type="pr_norm")
    return it[0], info


def study(name, A):
    n = A.shape[0]
    b = A @ np.random.default_rng(0).standard_normal(n)
    d = A.diagonal()
    out = {"matrix": name, "N": n, "nnz": int(A.nnz)}
    it, info = iterations(A, b, spla.LinearOperator(A.shape, lambda r: r / d), 100, 60)
    out["jacobi"] = it if info == 0 else ">%d" % it
    for R in (1, 2, 4, 8):
        Ab = A if R == 1 else block_diagonal(A, -(-n // R))
        Lm, Um = O.ilu0(Ab)
        it, info = iterations(A, b, spla.LinearOperator(A.shape, lambda r: O.apply(Lm, Um, r)), 100, 60)
        out["ilu0_R%d" % R] = it if info == 0 else ">%d" % it
    print(json.dumps(out), flush=True)


from oracle.cases.periodic_beam import PeriodicBeamCase
kw = dict(n=16, dt=1e-6, tf=1e-4, orderphi=4, ordereta=4, k=15)
K, Cm, M = [m.tocsr() for m in PeriodicBeamCase(**kw).assemble()]
A = (K + 0.5 / (0.25 * kw["dt"]) * Cm + 1.0 / (0.25 * kw["dt"] ** 2) * M).tocsr()
A.sort_indices()
study("5-1-1 Newmark matrix, n=16, order 4, dt=1e-6", A)
kw["dt"] = 1e-2
A = (K + 0.5 / (0.25 * kw["dt"]) * Cm + 1.0 / (0.25 * kw["dt"] ** 2) * M).tocsr()
A.sort_indices()
study("5-1-1 Newmark matrix, n=16, order 4, dt=1e-2", A)
m = 90
T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(m, m))
study("5-point grid operator 90 x 90 + 0.05 I", (sp.kron(sp.eye(m), T) + sp.kron(T, sp.eye(m)) + 0.05 * sp.eye(m * m)).tocsr())
