"""Runs the five BASELINE.json configurations at the paper's resolutions on one B200 and prints
one JSON line per configuration (sizes, assembly / factorisation / solve / step times, a result
checksum).  `gpurun -- python scripts/run_configs.py > gpurun_out/configs.jsonl`"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vlfs_b200
from vlfs_b200 import _lib as L

def emit(**kw):
    print(json.dumps(kw), flush=True)

def t(f):
    t0 = time.time(); r = f(); return r, time.time() - t0

# 5-1-1 periodic beam: n = 32, order 4, 100 steps (scripts/5-1-1-...jl:43-64)
from vlfs_b200.cases.periodic_beam import PeriodicBeamProblem, Periodic_Beam_params, run_periodic_beam
for n, order in ((32, 2), (64, 4)):
    p = Periodic_Beam_params(n=n, dt=1e-6, tf=1e-4, orderphi=order, ordereta=order, k=15)
    prob, ts = t(lambda: PeriodicBeamProblem(p))
    _, ta = t(prob.assemble)
    (u, v, a, states, st), tsolve = t(prob.solve)
    out, tall = t(lambda: run_periodic_beam(p))
    emit(config="5-1-1", n=n, order=order, N=prob.X.ndofs, nnz=prob.nnz, host_setup_s=ts, assemble_s=ta,
         newmark_100_steps_s=tsolve, ms_per_step=st["solve_ms"] / 100, e_phi=float(out[0][-1]), e_eta=float(out[1][-1]),
         run_periodic_beam_s=tall)

# 5-2-1 / 5-2-2 Khabakhpasheva, nx=20, ny=5, order=4 (scripts/5-2-1-...jl:31-35, 5-2-2-...jl)
from vlfs_b200.cases.khabakhpasheva import (KhabakhpashevaFreqProblem, KhabakhpashevaTimeProblem,
                                            Khabakhpasheva_freq_domain_params, Khabakhpasheva_time_domain_params)
for xi in (0.0, 625.0):
    prob, ts = t(lambda: KhabakhpashevaFreqProblem(Khabakhpasheva_freq_domain_params(xi=xi)))
    _, ta = t(prob.assemble)
    _, tf = t(lambda: prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12))
    (x, st), tsol = t(lambda: prob.sys.solve(vec=0))
    xs, eta = prob.postprocess(x)
    emit(config="5-2-1", xi=xi, N=prob.X.ndofs, nnz=prob.nnz, host_setup_s=ts, assemble_s=ta, factor_s=tf, solve_s=tsol,
         iterations=st["iterations"], rel_residual=st["rel_residual"], eta_max=float(eta.max()), eta_mean=float(eta.mean()))
prob, ts = t(lambda: KhabakhpashevaTimeProblem(Khabakhpasheva_time_domain_params(xi=0.0)))
_, ta = t(prob.assemble)
(tsv, hist, uva, st), trun = t(lambda: prob.solve(nsteps=2000))
xs, eta_rel = prob.postprocess(hist)
emit(config="5-2-2", N=prob.X.ndofs, nnz=prob.nnz, host_setup_s=ts, assemble_s=ta, newmark_2000_steps_s=trun,
     ms_per_step=st["solve_ms"] / 2000, envelope_max=float(eta_rel[1000:2000].max()), iterations=st["iterations"])

# 5-4-1 Yago, paper resolution, three wave lengths reusing pattern + symbolic LU
from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
prob, ts = t(lambda: YagoProblem(Yago_freq_domain_params(nx=32, ny=4, nz=3, order=4, lfactor=0.4, dfactor=4)))
first = True
for lf in (0.4, 0.6, 0.8):
    _, tw = t(lambda: prob.set_frequency(lf))
    _, ta = t(prob.assemble)
    if first:
        _, tf = t(lambda: prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=20, maxit=40, rtol=1e-10))
        first = False
    else:
        _, tf = t(lambda: prob.sys.refactor(0))
    (x, st), tsol = t(lambda: prob.sys.solve(vec=0))
    xs, rao = prob.rao(x)
    emit(config="5-4-1", lfactor=lf, N=prob.X.ndofs, nnz=prob.nnz, host_setup_s=ts, set_frequency_s=tw, assemble_s=ta,
         factor_s=tf, solve_s=tsol, iterations=st["iterations"], rel_residual=st["rel_residual"],
         rao_max=float(rao.max()), rao_mean=float(rao.mean()))

# 5-5-1 MultiGeo: coarse JSON order 2 (warm-up) and the fine gmsh mesh order 4
from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for f, order in ((os.path.join(ROOT, "tests/golden/multi_geo_coarse.json"), 2),
                 (os.environ.get("MULTIGEO_FINE", os.path.join(ROOT, "tests/golden/multi_geo.msh")), 4)):
    if not os.path.exists(f):
        emit(config="5-5-1", mesh=os.path.basename(f), skipped="mesh file not present")
        continue
    prob, ts = t(lambda: MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=f, order=order)))
    _, ta = t(prob.assemble)
    _, tf = t(lambda: prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-10))
    (x, st), tsol = t(lambda: prob.sys.solve(vec=0))
    eta = prob.X.split(x)[2]
    emit(config="5-5-1", mesh=os.path.basename(f), order=order, N=prob.X.ndofs, nnz=prob.nnz, host_setup_s=ts,
         assemble_s=ta, factor_s=tf, solve_s=tsol, iterations=st["iterations"], rel_residual=st["rel_residual"],
         eta_absmax=float(np.abs(eta).max()))

# 5-3-1 Liu: the reference's triangular model, ω = 0.4 and 0.8 (scripts/5-3-1-Liu.jl:20-32)
from vlfs_b200.cases.liu import LiuProblem, Liu_params
for om in (0.4, 0.8):
    prob, ts = t(lambda: LiuProblem(Liu_params(omega=om, mesh_file=os.path.join(ROOT, "tests/golden/Liu.json"))))
    _, ta = t(prob.assemble)
    _, tf = t(lambda: prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12))
    (x, st), tsol = t(lambda: prob.sys.solve(vec=0))
    xs, eta = prob.postprocess(x)
    emit(config="5-3-1", omega=om, N=prob.X.ndofs, nnz=prob.nnz, host_setup_s=ts, assemble_s=ta, factor_s=tf, solve_s=tsol,
         iterations=st["iterations"], rel_residual=st["rel_residual"], eta_max=float(eta.max()), eta_mean=float(eta.mean()))

# 5-1-4 periodic beam with free surface: n = 64, order 4, one wave period at T/40
# (scripts/5-1-4-periodic-beam-free-surface-energy.jl:73-86 uses n = 128, dt = T/200... scaled down)
from vlfs_b200.cases.periodic_beam_fs import PeriodicBeamFSProblem, Periodic_Beam_FS_params, run_periodic_beam_FS
kfs = 15
Tfs = 2 * np.pi / np.sqrt(9.81 * kfs * np.tanh(kfs * 1.0))
p = Periodic_Beam_FS_params(n=64, dt=Tfs / 40, tf=Tfs, order=4, k=kfs)
prob, ts = t(lambda: PeriodicBeamFSProblem(p))
_, ta = t(prob.assemble)
(u, v, a, states, st), tsolve = t(prob.solve)
out, tall = t(lambda: run_periodic_beam_FS(p))
emit(config="5-1-4", n=64, order=4, N=prob.X.ndofs, nnz=prob.nnz, host_setup_s=ts, assemble_s=ta,
     newmark_steps=int(states.shape[0]), newmark_s=tsolve, ms_per_step=st["solve_ms"] / max(states.shape[0], 1),
     e_phi=float(out[0][-1]), e_eta=float(out[1][-1]), E_kin_f_rel=float(out[2][-1] / out[6]),
     E_pot_f_rel=float(out[3][-1] / out[8]), run_periodic_beam_FS_s=tall)

# M2: the fine MultiGeo mesh refined uniformly, order 4 - 305 016 tets, N = 736 382
# (skipped unless RUN_M2=1: the factorisation of this 3-D mesh needs tens of GB)
if os.environ.get("RUN_M2") == "1" and os.path.exists(os.path.join(ROOT, "tests/golden/multi_geo.msh")):
    prob, ts = t(lambda: MultiGeoProblem(MultiGeo_freq_domain_params(
        mesh_file=os.path.join(ROOT, "tests/golden/multi_geo.msh"), order=4, refine=1)))
    _, ta = t(prob.assemble)
    _, ta2 = t(prob.assemble)
    rec = dict(config="M2", mesh="multi_geo.msh refined once", order=4, N=prob.X.ndofs, nnz=prob.nnz, host_setup_s=ts,
               assemble_first_s=ta, assemble_s=ta2)
    try:
        _, tf = t(lambda: prob.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-10))
        (x, st), tsol = t(lambda: prob.sys.solve(vec=0))
        rec.update(factor_s=tf, solve_s=tsol, iterations=st["iterations"], rel_residual=st["rel_residual"])
    except Exception as e:
        rec.update(solve_error=repr(e))
    emit(**rec)
