"""CPU preconditioner study on oracle matrices: does Jacobi / ILU Krylov
reach 1e-8 on the monolithic systems?  Numbers quoted in DESIGN.md §6."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from oracle.cases.yago import YagoCase
from oracle.cases.khabakhpasheva import KhabakhpashevaFreqCase

def study(name, A, b, maxit=2000):
    A = A.tocsc(); n = A.shape[0]
    xr = spla.splu(A).solve(b)
    def run(tag, M, solver):
        its = [0]
        def cb(_): its[0] += 1
        t0 = time.time()
        kw = dict(rtol=1e-10, maxiter=maxit, M=M, callback=cb)
        if solver is spla.gmres: kw.update(restart=100, callback_type='pr_norm', maxiter=maxit // 100)
        x, info = solver(A, b, **kw)
        err = np.linalg.norm(x - xr) / np.linalg.norm(xr)
        print('%-14s %-22s its %5d info %5d rel.err %.2e  %.1fs' % (name, tag, its[0], info, err, time.time() - t0), flush=True)
    d = A.diagonal(); d[d == 0] = 1
    J = spla.LinearOperator((n, n), matvec=lambda v: v / d, dtype=A.dtype)
    run('jacobi gmres(100)', J, spla.gmres)
    run('jacobi bicgstab', J, spla.bicgstab)
    for dt, ff in ((1e-4, 10), (1e-6, 40)):
        try:
            ilu = spla.spilu(A, drop_tol=dt, fill_factor=ff)
            M = spla.LinearOperator((n, n), matvec=ilu.solve, dtype=A.dtype)
            run('spilu(%g,%d) gmres' % (dt, ff), M, spla.gmres)
        except Exception as e:
            print(name, 'spilu failed', e)

c = YagoCase(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3); A, b = c.assemble(); study('Yago Y0', A, b)
c = KhabakhpashevaFreqCase(nx=10, ny=2, order=4, xi=0.0); A, b = c.assemble(); study('Khab K0', A, b)
