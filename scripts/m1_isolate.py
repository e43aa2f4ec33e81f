"""Diagnostic: one bilinear term of the fine MultiGeo mesh at a time, GPU against the compiled CPU interpreter
of the SAME descriptors."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.c_assembly import CSystem
from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
f = os.path.join(ROOT, "tests", "golden", "multi_geo.msh")
order = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pg = MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=f, order=order), device=0)
pc = MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=f, order=order), system_cls=CSystem)
nterms = [len(d.terms) for d in pc.sys.domains]
coefs = [[t["coef"] for t in d.terms] for d in pc.sys.domains]
print("terms per domain", nterms)
for di in (2, 3):
    for ti in range(nterms[di]):
        for dj, d in enumerate(pc.sys.domains):
            for tj, t in enumerate(d.terms):
                c = coefs[dj][tj] if (dj, tj) == (di, ti) else 0.0
                t["coef"] = complex(c)
                pg.sys.domains[dj].set_coef(tj, c)
        pg.sys.assemble(); v = pg.sys.get_values(0)
        mats, _ = pc.sys.assemble(); M = mats[0].tocsr(); M.sort_indices()
        dmax = np.abs(v - M.data).max(); ref = np.abs(M.data).max()
        nb = int((np.abs(v - M.data) > 1e-10 * ref).sum())
        print("domain", di, "term", ti, "max|A|", ref, "maxdiff", dmax, "nbad", nb, "nnz touched", int((np.abs(M.data) > 0).sum()))
