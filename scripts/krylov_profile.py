"""Krylov preconditioners and the basis-matrix sweep at the paper's resolutions on one B200 (evidence for
north_star (5) and SURVEY.md §8f): one JSON line per measurement.
`gpurun -- 'python scripts/krylov_profile.py > gpurun_out/krylov_profile.jsonl'`"""
import os, sys, time, json, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vlfs_b200
from vlfs_b200 import _lib as L


def emit(**kw):
    print(json.dumps(kw), flush=True)


def krylov_table(tag, S, mat, n, cplx, combos, maxit=3000, restart=100):
    rng = np.random.default_rng(0)
    xt = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    b = S.spmv(xt, mat)
    for combo in combos:
        mname, method, pname, precond = combo[:4]
        B = combo[4] if len(combo) > 4 else 0
        if BLOCKS_ONLY and not B:
            continue
        if B:
            pname = "%s(block_rows=%d)" % (pname, B)
        try:
            t0 = time.perf_counter()
            S.solver_setup(mat, method=method, precond=precond, restart=restart, maxit=maxit, rtol=1e-10, block_size_hint=B)
            t_setup = time.perf_counter() - t0
            x, st = S.solve(b=b)
            rec = dict(case=tag, method=mname, precond=pname, N=n, nnz=S.nnz, setup_s=t_setup, iterations=st["iterations"],
                       converged=int(st["converged"]), rel_residual=st["rel_residual"], solve_ms=st["solve_ms"],
                       error=float(np.linalg.norm(x - xt) / np.linalg.norm(xt)))
            if precond == L.PRECOND_BLOCK_ILU0:
                S.precond_apply(b)
                rec["ilu0"] = S.precond_info()
            emit(**rec)
        except Exception as e:
            emit(case=tag, method=mname, precond=pname, error=repr(e))


BLOCKS_ONLY = os.environ.get("KRYLOV_PROFILE") == "blocks"     # second pass: only the block-Jacobi ILU(0) variants
GM, BI, CG_ = ("gmres", L.SOLVER_GMRES), ("bicgstab", L.SOLVER_BICGSTAB), ("cg", L.SOLVER_CG)
ILU, JAC, LU = ("ilu0", L.PRECOND_BLOCK_ILU0), ("jacobi", L.PRECOND_JACOBI), ("sparse_lu", L.PRECOND_SPARSE_LU)

try:    # 5-1-1 periodic beam, n = 64, order 4: Newmark matrix K + γ/(βΔt) C + 1/(βΔt²) M
    from vlfs_b200.cases.periodic_beam import PeriodicBeamProblem, Periodic_Beam_params, MAT_A
    prob = PeriodicBeamProblem(Periodic_Beam_params(n=64, dt=1e-6, tf=1e-4, orderphi=4, ordereta=4, k=15))
    prob.assemble()
    krylov_table("5-1-1 n=64 order 4 (Newmark matrix)", prob.sys, MAT_A, prob.X.ndofs, False,
                 [GM + ILU, BI + ILU, GM + JAC, GM + LU, GM + ILU + (1024,), BI + ILU + (1024,), GM + ILU + (256,)])
    del prob
except Exception:
    emit(case="5-1-1", error=traceback.format_exc())

try:    # 5-2-2 Khabakhpasheva time domain, nx = 20, ny = 5, order 4
    from vlfs_b200.cases.khabakhpasheva import KhabakhpashevaTimeProblem, Khabakhpasheva_time_domain_params
    from vlfs_b200.cases import khabakhpasheva as K
    prob = KhabakhpashevaTimeProblem(Khabakhpasheva_time_domain_params(xi=0.0))
    prob.assemble()
    krylov_table("5-2-2 nx=20 ny=5 order 4 (Newmark matrix)", prob.sys, K.MAT_A, prob.X.ndofs, False,
                 [GM + ILU, BI + ILU, GM + JAC, GM + LU, BI + ILU + (1024,)])
    del prob
except Exception:
    emit(case="5-2-2", error=traceback.format_exc())

try:    # 5-4-1 Yago at the paper's resolution (Y1): the indefinite frequency-domain operator
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    prob = YagoProblem(Yago_freq_domain_params(nx=32, ny=4, nz=3, order=4, lfactor=0.4, dfactor=4), basis=True)
    t0 = time.perf_counter(); prob.assemble_basis(); t_basis = time.perf_counter() - t0
    ms_asm = prob.sys.assemble_device_ms(reps=5)          # four matrices at once on the union pattern
    ts = []
    for lf in (0.4, 0.6, 0.8, 0.4, 0.6, 0.8):
        prob.set_frequency(lf)
        t0 = time.perf_counter(); prob.combine_frequency(); ts.append((time.perf_counter() - t0) * 1e3)
    if not BLOCKS_ONLY: emit(case="Y1 basis sweep", N=prob.X.ndofs, nnz=prob.nnz, assemble_basis_first_s=t_basis, assemble_4_basis_matrices_ms=ms_asm,
         combine_ms=ts, note="combine = one vlfs_combine incl. host call + stream sync; re-assembly of A(w) alone: bench.py ms_per_step")
    krylov_table("Y1 A(w), lfactor 0.8", prob.sys, prob.MAT_A, prob.X.ndofs, True, [GM + ILU, GM + JAC, GM + ILU + (1024,)],
                 maxit=200, restart=50)
    del prob
except Exception:
    emit(case="Y1", error=traceback.format_exc())
