"""Diagnostic: GPU matrix of the fine MultiGeo mesh against the oracle, per field block."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("VLFS_PKG_ROOT"):
    sys.path.insert(0, os.environ["VLFS_PKG_ROOT"])
from oracle.cases.multigeo import MultiGeoCase
from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
f = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", sys.argv[1] if len(sys.argv) > 1 else "multi_geo.msh")
order = int(sys.argv[2]) if len(sys.argv) > 2 else 4
prob = MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=f, order=order), device=0)
prob.assemble()
rp, ci = prob.sys.get_pattern(); v = prob.sys.get_values(0)
ref = MultiGeoCase(f, order=order); A, b = ref.assemble(); A = A.tocsr(); A.sort_indices()
print("pattern", np.array_equal(rp, A.indptr), np.array_equal(ci, A.indices))
d = np.abs(v - A.data)
rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
o = prob.X.offsets
for fi in range(len(o) - 1):
    for fj in range(len(o) - 1):
        m = (rows >= o[fi]) & (rows < o[fi + 1]) & (ci >= o[fj]) & (ci < o[fj + 1])
        if m.any():
            print(fi, fj, "max|A|", np.abs(A.data[m]).max(), "maxdiff", d[m].max(), "nbad", int((d[m] > 1e-10 * np.abs(A.data[m]).max()).sum()), "of", int(m.sum()))
bad = np.where(d > 1e-10 * np.abs(A.data).max())[0]
br = np.unique(rows[bad])
print("bad rows", len(br), br[:10], br[-10:])
np.savez("gpurun_out/m1_bad.npz", bad=bad, rows=rows[bad], cols=ci[bad], got=v[bad], want=A.data[bad])
for dname in ("dom_Lb", "dom_Gb", "Lb", "Gb"):
    if hasattr(prob, dname):
        print(dname, type(getattr(prob, dname)))
print("domains", [(getattr(dd, "ncells", None)) for dd in getattr(prob.sys, "domains", [])])
