"""CPU study (oracle matrices): GMRES preconditioned by block-Jacobi with exact subdomain LU on
x-slab row partitions - the multi-GPU solve design of SURVEY.md §8e.  Iterations to 1e-10."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from oracle.cases.yago import YagoCase
from oracle.fem import FESpace

def dof_x(case):
    xs = []
    for V in (case.V_phi, case.V_kap, case.V_eta):
        xs.append(V.dof_coords()[:, 0])
    return np.concatenate(xs)

def study(case, parts):
    A, b = case.assemble(); A = A.tocsr(); n = A.shape[0]
    x = dof_x(case)
    xr = spla.splu(A.tocsc()).solve(b)
    for P in parts:
        edges = np.quantile(x, np.linspace(0, 1, P + 1)); edges[-1] += 1
        owner = np.clip(np.searchsorted(edges, x, side='right') - 1, 0, P - 1)
        for overlap in (0, 1):
            lus, idxs, cores = [], [], []
            for r in range(P):
                own = np.nonzero(owner == r)[0]
                idx = own
                for _ in range(overlap):     # algebraic overlap: add neighbours in the graph
                    idx = np.unique(np.concatenate([idx, A[idx].indices]))
                lus.append(spla.splu(A[idx][:, idx].tocsc())); idxs.append(idx)
                cores.append(np.isin(idx, own))
            def M(v):
                z = np.zeros_like(v)
                for lu, idx, core in zip(lus, idxs, cores):
                    s = lu.solve(v[idx])
                    z[idx[core]] = s[core]          # restricted additive Schwarz
                return z
            its = [0]
            def cb(_): its[0] += 1
            xs_, info = spla.gmres(A, b, rtol=1e-10, restart=200, maxiter=3, M=spla.LinearOperator((n, n), matvec=M, dtype=A.dtype), callback=cb, callback_type='pr_norm')
            print('P=%d overlap=%d its=%d info=%d err=%.2e' % (P, overlap, its[0], info, np.linalg.norm(xs_ - xr) / np.linalg.norm(xr)), flush=True)

print('Y0 warm-up'); study(YagoCase(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3), (2, 4, 8))
print('nx=4 ny=1 nz=2 order 4'); study(YagoCase(nx=4, ny=1, nz=2, order=4, lfactor=0.4, dfactor=4), (2, 4, 8))
