import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
import vlfs_b200
from vlfs_b200 import _lib as L
from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
for kw in [dict(nx=8,ny=2,nz=2), dict(nx=16,ny=2,nz=3), dict(nx=16,ny=4,nz=3), dict(nx=32,ny=2,nz=3), dict(nx=32,ny=4,nz=2), dict(nx=32,ny=4,nz=3)]:
    for leaf in (96, 32):
        p=YagoProblem(Yago_freq_domain_params(order=4,lfactor=0.4,dfactor=4,**kw)); p.assemble()
        S=p.sys
        S.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=1, maxit=1, rtol=1e-14, block_size_hint=leaf)
        x,st=S.solve(vec=0)
        info=S.direct_info()
        S.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=30, rtol=1e-11, block_size_hint=leaf)
        x,st2=S.solve(vec=0)
        print(kw, 'leaf',leaf,'N',S.ndofs,'res1',st['rel_residual'],'its',st2['iterations'],st2['rel_residual'],'maxfront',info['maxfront'],'levels',info['nlevels'],'num_ms',info['numeric_ms'], flush=True)
        S.close()
