"""Diagnostic: residual history of the Y1 solve with and without the stagnation stop (VLFS_DEBUG=1 prints it)."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlfs_b200
from vlfs_b200 import _lib as L
from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
p = YagoProblem(Yago_freq_domain_params(nx=nx, ny=4, nz=3, order=4, lfactor=0.4, dfactor=4), device=0)
p.assemble()
p.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=4, maxit=16, rtol=1e-15)
x, st = p.sys.solve(vec=0)
print(json.dumps({"nx": nx, "N": p.sys.ndofs, "stats": {k: (float(v) if not isinstance(v, (bool, int)) else v) for k, v in st.items()}}))
A = p.sys.get_matrix(0); b = p.sys.get_vector(0)
r = b - A @ x
absA = abs(A)
print("host residual", np.linalg.norm(r) / np.linalg.norm(b), "eps|A||x|/|b|", np.finfo(float).eps * np.linalg.norm(absA @ abs(x)) / np.linalg.norm(b),
      "max|A|", abs(A.data).max(), "|x|/|b|", np.linalg.norm(x) / np.linalg.norm(b))
