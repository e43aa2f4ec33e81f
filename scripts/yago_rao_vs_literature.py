"""RAO of the Yago plate (configuration 5-4-1) against the curves the reference plots it with
(scripts/5-4-1-Yago.jl:73-127): Fu et al. (numerical, `data/Ref_data/Fu/L_0x.csv`, abscissa
-2(x-0.5)) and Yago & Endo (experiment, `data/Ref_data/Yago/lambda_L_0.x.csv`).  Copies of the six
files are under tests/golden/.

  gpurun -- python scripts/yago_rao_vs_literature.py                 # GPU, paper resolution
  python scripts/yago_rao_vs_literature.py --cpu --nx 8 --ny 2 --nz 3   # CPU oracle path (SuperLU), minutes

One JSON line per wave length: max / mean |RAO - curve| at the curve's abscissae."""
import argparse
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true", help="compiled CPU restatement + SuperLU instead of the GPU")
    ap.add_argument("--nx", type=int, default=32)
    ap.add_argument("--ny", type=int, default=4)
    ap.add_argument("--nz", type=int, default=3)
    args = ap.parse_args()
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    from vlfs_b200 import _lib as L
    prob = None
    for lf, fu, ya in ((0.4, "L_04", "0.4"), (0.6, "L_06", "0.6"), (0.8, "L_08", "0.8")):
        par = Yago_freq_domain_params(nx=args.nx, ny=args.ny, nz=args.nz, order=4, lfactor=lf, dfactor=4)
        if args.cpu:
            import scipy.sparse.linalg as spla
            from oracle.c_assembly import CSystem
            prob = YagoProblem(par, system_cls=CSystem)
            mats, vecs = prob.sys.assemble()
            x = spla.splu(mats[0].tocsc()).solve(vecs[0])
        else:
            if prob is None:
                prob = YagoProblem(par)
                first = True
            prob.set_frequency(lf)
            prob.assemble()
            if first:
                prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=20, maxit=40, rtol=1e-10)
                first = False
            else:
                prob.sys.refactor(0)
            x, st = prob.sys.solve(vec=0)
        xs, rao = prob.rao(x)
        o = np.argsort(xs)
        xu, iu = np.unique(xs[o], return_index=True)
        ru = rao[o][iu]
        F = np.loadtxt(os.path.join(GOLD, "Fu_%s.csv" % fu), delimiter=",")
        Y = np.loadtxt(os.path.join(GOLD, "Yago_lambda_L_%s.csv" % ya), delimiter=",")
        dF = np.abs(np.interp(-2.0 * (F[:, 0] - 0.5), xu, ru) - F[:, 1])
        dY = np.abs(np.interp(Y[:, 0], xu, ru) - Y[:, 1])
        print(json.dumps(dict(lfactor=lf, N=int(prob.X.ndofs), path="cpu" if args.cpu else "gpu",
                              vs_Fu_max=float(dF.max()), vs_Fu_mean=float(dF.mean()),
                              vs_Yago_experiment_max=float(dY.max()), vs_Yago_experiment_mean=float(dY.mean()),
                              rao_min=float(ru.min()), rao_max=float(ru.max()))), flush=True)


if __name__ == "__main__":
    main()
