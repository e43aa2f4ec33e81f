"""One numeric refactorisation + one solve of the Y1 multifrontal LU, for `ncu` launch lists:
  ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file lu.csv python scripts/lu_profile.py
Prints the front-size histogram and the per-level numbers the host analysis produced."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from vlfs_b200 import _lib as L
from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params

kw = dict(nx=32, ny=4, nz=3, order=4, lfactor=0.4, dfactor=4)
if len(sys.argv) > 1 and sys.argv[1] == "small":
    kw = dict(nx=8, ny=2, nz=3, order=4, lfactor=0.4, dfactor=4)
p = YagoProblem(Yago_freq_domain_params(**kw))
p.assemble()
S = p.sys
t = time.time()
S.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=20, maxit=40, rtol=1e-10)
print("setup s", time.time() - t, S.direct_info(), flush=True)
t = time.time()
S.refactor(0)
x, st = S.solve(vec=0)
print("refactor+solve s", time.time() - t, S.direct_info()["numeric_ms"], st, flush=True)
