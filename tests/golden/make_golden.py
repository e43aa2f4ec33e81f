"""Generates the committed oracle fixtures `oracle_*.npz` (run from the repository root:
`python tests/golden/make_golden.py`).  They freeze the oracle's outputs for the warm-up
configurations of the reference scripts so that (i) the CPU suite detects any drift of the
oracle itself and (ii) the GPU suite has golden vectors that do not depend on re-running the
oracle.  The reference cannot be executed here (no Julia), so these are oracle
outputs, not Gridap outputs - see DESIGN.md §2 for what pins the oracle."""
import os
import sys
import hashlib
import numpy as np
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def pattern_digest(A):
    A = A.tocsr(); A.sort_indices()
    h = hashlib.sha256()
    h.update(np.asarray(A.indptr, dtype=np.int64).tobytes())
    h.update(np.asarray(A.indices, dtype=np.int64).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def probe(A, b):
    """Pattern digest + a handful of size-independent value functionals."""
    A = A.tocsr(); A.sort_indices()
    n = A.shape[0]
    x = np.sin(0.1 * np.arange(n)) + (1j * np.cos(0.1 * np.arange(n)) if np.iscomplexobj(A.data) else 0)
    return dict(n=n, nnz=A.nnz, digest=pattern_digest(A), Ax=A @ x, diag=A.diagonal(), b=b,
                absmax=np.abs(A.data).max(), sum=A.data.sum())


def main():
    from oracle.cases.yago import YagoCase
    from oracle.cases.khabakhpasheva import KhabakhpashevaFreqCase, KhabakhpashevaTimeCase
    from oracle.cases.periodic_beam import PeriodicBeamCase
    from oracle.cases.multigeo import MultiGeoCase
    # 5-4-1 warm-up (scripts/5-4-1-Yago.jl:17-30)
    c = YagoCase(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3)
    A, b = c.assemble()
    x = spla.splu(A.tocsc()).solve(b)
    xs, rao = c.rao(x)
    np.savez_compressed(os.path.join(HERE, "oracle_yago_warmup.npz"), x=x, rao_x=xs, rao=rao, **probe(A, b))
    # 5-2-1 warm-up (scripts/5-2-1-...jl:17-28)
    c = KhabakhpashevaFreqCase(nx=10, ny=1, order=2, xi=0.0)
    A, b = c.assemble()
    x = spla.splu(A.tocsc()).solve(b)
    xs, eta = c.eta_postprocess(x)
    np.savez_compressed(os.path.join(HERE, "oracle_khab_freq_warmup.npz"), x=x, xs=xs, eta=eta, **probe(A, b))
    # 5-2-2 warm-up, first 40 steps
    c = KhabakhpashevaTimeCase(nx=10, ny=1, order=2, xi=0.0)
    ts, hist, (u, v, a) = c.run(nsteps=40)
    np.savez_compressed(os.path.join(HERE, "oracle_khab_time_warmup.npz"), ts=ts, eta_hist=hist, u=u)
    # 5-1-1 warm-up like (scripts/5-1-1-...jl:20-40: n=3, order 2, k=1) and one convergence point
    c = PeriodicBeamCase(n=3, dt=0.5, tf=1.0, orderphi=2, ordereta=2, k=1)
    r = c.run()
    c2 = PeriodicBeamCase(n=8, dt=1e-6, tf=1e-4, orderphi=2, ordereta=2, k=15)
    r2 = c2.run()
    np.savez_compressed(os.path.join(HERE, "oracle_periodic_beam.npz"), u=r["u"], e_phi=r["e_phi"], e_eta=r["e_eta"],
                        E=np.array(r["E"]), conv_e_phi=r2["e_phi"][-1], conv_e_eta=r2["e_eta"][-1])
    # 5-5-1 warm-up (scripts/5-5-1-MultiGeo.jl:16-25)
    c = MultiGeoCase(os.path.join(HERE, "multi_geo_coarse.json"), order=2)
    A, b = c.assemble()
    x = spla.splu(A.tocsc()).solve(b)
    np.savez_compressed(os.path.join(HERE, "oracle_multigeo_warmup.npz"), x=x, **probe(A, b))
    # 5-3-1 Liu, ω = 0.4 (scripts/5-3-1-Liu.jl:20-24) on the reference's model file; the value
    # functionals are kept at every 16th entry to keep the fixture small
    from oracle.cases.liu import LiuCase
    from oracle.cases.periodic_beam_fs import PeriodicBeamFSCase
    c = LiuCase(os.path.join(HERE, "Liu.json"), omega=0.4)
    A, b = c.assemble()
    x = spla.splu(A.tocsc()).solve(b)
    xs, eta = c.postprocess(x)
    pr = probe(A, b)
    np.savez_compressed(os.path.join(HERE, "oracle_liu_omega04.npz"), n=pr["n"], nnz=pr["nnz"], digest=pr["digest"],
                        Ax=pr["Ax"][::16], diag=pr["diag"][::16], b=pr["b"][::16], xs=xs, eta=eta,
                        eta_dofs=x[c.X.offsets[2]:c.X.offsets[3]])
    # 5-1-4 on a coarse mesh: the reference's tuple of errors and energies
    r = PeriodicBeamFSCase(n=4, dt=0.01, tf=0.05, order=2, k=1).run()
    np.savez_compressed(os.path.join(HERE, "oracle_periodic_beam_fs.npz"), u=r["u"], e_phi=r["e_phi"],
                        e_eta=r["e_eta"], E=np.array(r["E"]))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
