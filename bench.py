#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 hot path (assembly nnz/s + SpMV GB/s + solve
s/step on the Yago 3-D case, BASELINE.json) with the CPU restatement timed beside it.

  python bench.py --gpus N --steps K --warmup W [--workload Y1]        (our arm)
  python bench.py --impl reference --gpus N --steps K --warmup W       (CPU reference arm)

A "step" is one numeric pass of `AffineFEOperator(a,b,X,Y)` for one frequency: zero + cell
integration of every term + scatter into the fixed CSR pattern (matrix and right-hand side).
`value` = nnz of the monolithic matrix / step time, inputs resident in HBM.  `e2e` is the
same metric through the host-buffer C-ABI path: coefficient fields copied host->device, the
assembly, and the matrix values + rhs copied device->host, all inside the timed region.
SpMV and solve are timed in the same run and reported in `spmv` / `solve`.
The CPU arm (`cpu_baseline`, `--impl reference`) times the compiled restatement of the same path
(oracle/c/asm_cpu.cpp, all host threads) on the same workload, plus SciPy's CSR SpMV and SuperLU
on an eighth-size sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Yago parameters, description)   [scripts/5-4-1-Yago.jl:17-46; SURVEY.md §8(d)]
    "Y0": dict(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3),
    "Ymid": dict(nx=8, ny=2, nz=2, order=4, lfactor=0.4, dfactor=4),
    "Y1": dict(nx=32, ny=4, nz=3, order=4, lfactor=0.4, dfactor=4),
    "Y2": dict(nx=64, ny=8, nz=6, order=4, lfactor=0.4, dfactor=4),
}


_real_stdout = None


def emit(obj):
    f = _real_stdout or sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel from the committed
    `ncu --set full` capture (profiles/r1_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The timed region of this
    workload is a few milliseconds, so the sampler polls NVML in a thread (≈ 0.3 ms period)
    instead of `nvidia-smi -lms` (100 ms); nvidia-smi stays as the fallback."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.rows = []          # (sm_mhz, reasons bitmask)
        self.proc = None
        self.nv = None
        self.run = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        try:
            mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                why = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                why = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            self.rows.append((mhz, why))
        except Exception:
            pass

    def _poll(self):
        while self.run:
            self._sample()
            time.sleep(0.0003)

    def start(self):
        if self.nv is not None:
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self._sample()          # one sample right at the end of the timed region, whatever its length
            self.run = False
            self.thread.join(timeout=1.0)
            nv = self.nv
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            sm = sorted(r[0] for r in self.rows)
            reasons = sorted(nm for nm, bit in bits.items() if any(r[1] & bit for r in self.rows))
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                    "samples": len(sm), "sampler": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for k, nm in enumerate(self.NAMES):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampler": "nvidia-smi"}


def assembly_bytes(prob, nnz, ncoo, sT):
    """Algorithmic bytes of one numeric assembly pass: every CSR value written
    once, the COO->slot permutation read once, DOF maps / coordinates / coefficient fields
    read once."""
    tables = 0
    for t, nd, nv, D in [(prob.Om, 27, 8, 3), (prob.Gf, 27 + prob.V_Gk.reffe.ndofs, 8 + 4, 3),
                         (prob.Gb, 27 + prob.V_Ge.reffe.ndofs, 8 + 4, 3)]:
        tables += t.ncells * (4 * nd + 8 * D * nv)
    tables += prob.Gf.ncells * 25 * 8 * 5
    return nnz * sT + 4 * ncoo + tables


def run_ours(args):
    import numpy as np
    import torch
    import vlfs_b200
    from vlfs_b200 import _lib as L
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kw = WORKLOADS[args.workload]
    t0 = time.time()
    prob = YagoProblem(Yago_freq_domain_params(**kw), device=local)
    setup_s = time.time() - t0
    S = prob.sys
    nnz, ncoo, N = S.nnz, S.ncoo, S.ndofs
    sT = 16
    peak, peak_kind = measured_peaks()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxms(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    launches0 = S.launches()
    for _ in range(max(args.warmup, 3)):
        S.assemble()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    launches1 = S.launches()
    ms = maxms(S.assemble_device_ms(reps=args.steps))
    launches2 = S.launches()
    barrier()
    clk = clocks.stop()
    value = world * nnz / (ms * 1e-3)

    # per-kernel times (CUDA events on the launching stream) for the roofline of the dominant kernel
    prof = S.assemble_profile(reps=3)
    dom_name, dom_ms = max(prof, key=lambda kv: kv[1])
    rows_ms = sum(t for n, t in prof if n.startswith("assemble_rows"))
    if rows_ms > 0:
        # row-centric pass (one kernel, launched for the early and the late rows): every CSR value
        # written once + the COO->slot permutation read once = SURVEY.md §8(d)'s per-unit figure
        dom_name, dom_ms = "assemble_rows", rows_ms
        kbytes = nnz * sT + 4 * ncoo
    elif dom_name.startswith("gather_reference"):
        # slot-centric pass: every value written once, one index per COO entry, one run pointer per slot
        kbytes = nnz * sT + 4 * ncoo + 4 * nnz
    else:
        # cell-centric pass on Γf: the coefficient-field products (675 + 675 real, 729 imaginary
        # entries per cell) reduced into their slots, dest read once, per-cell tables read once
        nk = prob.V_Gk.reffe.ndofs
        kbytes = prob.Gf.ncells * ((nk * 27 * 2 + 27 * 27) * 8 + (27 + nk) ** 2 * 4
                                   + 12 * 24 + (27 + nk) * 4 + 5 * 25 * 8)
    roof = {"bound": "hbm", "kernel": dom_name, "achieved": kbytes / (dom_ms * 1e-3) / 1e9, "peak": peak,
            "peak_kind": peak_kind, "unit": "GB/s",
            "traffic": ncu_traffic(dom_name) if args.workload == "Y1" else None,   # the ncu capture is of Y1
            "share_of_step": dom_ms / sum(t for _, t in prof), "algorithmic_bytes": kbytes}
    roof["frac"] = roof["achieved"] / peak
    step_bytes = assembly_bytes(prob, nnz, ncoo, sT)

    # SpMV (device resident, x fixed = (sin 0.1 i, cos 0.1 i))
    S.spmv_device_ms(reps=3)
    barrier()
    sp_ms = maxms(S.spmv_device_ms(reps=max(args.steps, 10)))
    barrier()
    sp_bytes = nnz * (sT + 4) + 4 * (N + 1) + 2 * N * sT
    spmv = {"ms": sp_ms, "GBs": world * sp_bytes / (sp_ms * 1e-3) / 1e9, "bytes": sp_bytes,
            "traffic": ncu_traffic("csr_spmv") if args.workload == "Y1" else None}
    spmv["frac"] = spmv["GBs"] / world / peak

    # solve: sparse LU factorisation + GMRES(1-2 its) per frequency
    solve = None
    if not args.no_solve:
        t0 = time.time()
        S.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=20, maxit=40, rtol=1e-10)
        first_s = time.time() - t0
        info = S.direct_info()
        samples = []
        for _ in range(3):                      # one frequency step = refactorise + solve; median of 3
            t0 = time.time()
            S.refactor(0)
            x, st = S.solve(vec=0)
            samples.append(time.time() - t0)
        step_s = sorted(samples)[1]
        info = S.direct_info()
        xs, rao = prob.rao(x)
        solve = {"s_per_step": step_s, "s_per_step_samples": samples, "first_setup_s": first_s,
                 "symbolic_ms": info["symbolic_ms"],
                 "numeric_ms": info["numeric_ms"], "gmres_iterations": st["iterations"],
                 "rel_residual": st["rel_residual"], "solve_ms": st["solve_ms"],
                 "factor_TFLOPs": info["flops"] / (info["numeric_ms"] * 1e-3) / 1e12,
                 "fill_entries": info["fill"], "fronts": info["nfronts"], "max_front": info["maxfront"],
                 "rao_max": float(rao.max()) if len(rao) else None}

    # end to end through host buffers: H2D coefficient fields, assemble, D2H values + rhs
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h2d = 0
    vals_host = torch.empty(nnz * 2, dtype=torch.float64).pin_memory().numpy().view(np.complex128)
    rhs_host = torch.empty(N * 2, dtype=torch.float64).pin_memory().numpy().view(np.complex128)
    import ctypes as C

    parts = [0.0, 0.0, 0.0]
    # the step's inputs: coefficient fields at the quadrature points, evaluated on the host once
    # (the analytic lambdas of scripts/5-4-1-Yago.jl) and held in pinned memory
    inputs = prob.frequency_inputs(kw["lfactor"])
    for key in ("weights_Gf", "weights_Gin"):
        inputs[key] = [pin(a) for a in inputs[key]]

    def e2e_step():
        t_a = time.perf_counter()
        prob.apply_frequency_inputs(inputs)         # H2D of the ω-dependent fields (pinned host arrays)
        t_b = time.perf_counter()
        S.assemble()
        t_c = time.perf_counter()
        S._check(S.lib.vlfs_get_values(S.ctx, 0, vals_host.ctypes.data_as(C.c_void_p)))
        S._check(S.lib.vlfs_get_vector(S.ctx, 0, rhs_host.ctypes.data_as(C.c_void_p)))
        t_d = time.perf_counter()
        parts[0] += t_b - t_a; parts[1] += t_c - t_b; parts[2] += t_d - t_c
    h2d = (5 * prob.Gf.ncells * 25 + 2 * prob.Gin.ncells * 25) * 8
    d2h = nnz * sT + N * sT
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    nrep = max(2, min(args.steps, 5))
    parts[:] = [0.0, 0.0, 0.0]
    for _ in range(nrep):
        e2e_step()
    torch.cuda.synchronize()
    e_ms = maxms((time.perf_counter() - t0) * 1e3 / nrep)
    barrier()
    e2e = {"value": world * nnz / (e_ms * 1e-3), "unit": "nnz/s", "ms_per_step": e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "parts_ms": {"set_coefficients_h2d": parts[0] * 1e3 / nrep, "assemble": parts[1] * 1e3 / nrep,
                        "values_d2h": parts[2] * 1e3 / nrep}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline(kw, with_solve=not args.no_cpu_solve)
        except Exception as e:              # the CPU leg must not take the GPU numbers down with it
            cpu = {"error": repr(e), "kind": "port"}

    if rank == 0:
        out = {
            "metric": "assembly nnz/s (Yago 3D; + SpMV GB/s + solve s/step in `spmv`/`solve`)",
            "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (complex128 matrix)", "data": "synthetic (deterministic Yago mesh, analytic coefficients)",
            "config": {"workload": "%s: Yago 5-4-1 %s, N=%d, nnz=%d, ncoo=%d; inputs > L2 (values %.2f GB + dest %.2f GB), no flush"
                       % (args.workload, json.dumps(kw), N, nnz, ncoo, nnz * sT / 1e9, ncoo * 4 / 1e9),
                       "parallelism": "replicas" if world > 1 else "single"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches2 - launches1),
            "roofline": roof, "step_GBs": step_bytes / (ms * 1e-3) / 1e9, "step_frac": step_bytes / (ms * 1e-3) / 1e9 / peak,
            "kernels_ms": prof, "spmv": spmv, "solve": solve, "cpu_baseline": cpu,
            "host_setup_s": setup_s,
        }
        emit(out)
    if dist is not None:
        dist.destroy_process_group()


def run_ours_distributed(args):
    """N > 1: the Yago mesh refined N x along x (per-GPU work fixed = Y1: weak scaling), rows
    partitioned into N x-slabs, one rank per GPU.  Assembly needs no communication; SpMV does
    one halo exchange; the solve is distributed GMRES with per-rank multifrontal LU."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import vlfs_b200
    from vlfs_b200 import _lib as L, dist as D
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kw = dict(WORKLOADS[args.workload])
    kw["nx"] = kw["nx"] * world
    sT = 16
    peak, peak_kind = measured_peaks()
    t0 = time.time()
    mk = lambda *a, **k: D.DistributedSystem(*a, rank=rank, nranks=world, **{**k, "device": local})
    prob = YagoProblem(Yago_freq_domain_params(**kw), device=local, system_cls=mk)
    setup_s = time.time() - t0
    dsys = prob.sys
    S = dsys.local
    ops = D.GpuOps(dsys)
    rp, _ = S.get_pattern()
    nnz_owned = int(rp[dsys.n_owned])

    def allsum(v):
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def allmax(v):
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    nnz_global = int(allsum(nnz_owned))
    N_global = dsys.ndofs_global
    for _ in range(max(args.warmup, 3)):
        S.assemble()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    l1 = S.launches()
    ms = allmax(S.assemble_device_ms(reps=args.steps))
    l2 = S.launches()
    barrier()
    clk = clocks.stop()
    value = nnz_global / (ms * 1e-3)
    prof = S.assemble_profile(reps=3)
    dom_name, dom_ms = max(prof, key=lambda kv: kv[1])
    step_bytes_local = S.nnz * sT + 4 * S.ncoo
    roof = {"bound": "hbm", "kernel": dom_name + " (rank 0)", "achieved": step_bytes_local * (dom_ms / sum(t for _, t in prof)) / (dom_ms * 1e-3) / 1e9,
            "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "traffic": None,
            "share_of_step": dom_ms / sum(t for _, t in prof)}
    roof["frac"] = roof["achieved"] / peak

    # distributed SpMV: halo exchange (NCCL send/recv) + local product over the owned rows
    ops.bind_stream()
    member = D.Member(dsys, ops)
    fleet = D.Fleet([member], world, dist=dist)
    i = dsys.local_gids.astype(np.float64)
    xl = np.sin(0.1 * i) + 1j * np.cos(0.1 * i)
    xs, ys = [ops.from_host(xl)], [ops.zeros()]
    for _ in range(3):
        fleet.matvec(xs, ys)
    barrier()
    reps = max(args.steps, 10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fleet.matvec(xs, ys)
    e1.record()
    torch.cuda.synchronize()
    sp_ms = allmax(e0.elapsed_time(e1) / reps)
    barrier()
    sp_bytes = nnz_global * (sT + 4) + 4 * (N_global + 1) + 2 * N_global * sT
    spmv = {"ms": sp_ms, "GBs": sp_bytes / (sp_ms * 1e-3) / 1e9, "bytes": sp_bytes,
            "frac": sp_bytes / (sp_ms * 1e-3) / 1e9 / world / peak, "halo_bytes_rank0": fleet.halo_bytes}

    solve = None
    if not args.no_solve:
        # exact distributed solve: per-rank partial multifrontal LU (interface kept), dense interface
        # LU on rank 0, used as the preconditioner of distributed GMRES (1-2 refinement iterations)
        t0 = time.time()
        schur = D.SchurSolver(fleet)
        ok = schur.setup()
        torch.cuda.synchronize()
        first_s = allmax(time.time() - t0)
    if not args.no_solve and not ok:
        solve = {"error": "a rank could not factorise its subdomain: %s" % (schur.error or "see another rank")}
    elif not args.no_solve:
        info = S.direct_info()
        bs, x0 = [ops.rhs_vector()], [ops.zeros()]
        barrier()
        t0 = time.time()
        schur.refactor()
        torch.cuda.synchronize()
        factor_s = allmax(time.time() - t0)
        t1 = time.time()
        st = fleet.gmres(bs, x0, restart=8, maxit=8, rtol=1e-10, apply=schur.apply, stagnation=0.25)
        torch.cuda.synchronize()
        gm_s = allmax(time.time() - t1)
        step_s = allmax(time.time() - t0)
        solve = {"s_per_step": step_s, "first_setup_s": first_s, "refactor_s": factor_s, "gmres_s": gm_s,
                 "symbolic_ms": info["symbolic_ms"], "numeric_ms": info["numeric_ms"],
                 "gmres_iterations": st["iterations"], "rel_residual": st["rel_residual"],
                 "converged": st["converged"], "stagnated_at_attainable_accuracy": st["stagnated"],
                 "interface_unknowns": int(schur.ng), "interface_solver": schur.interface_solver,
                 "method": "substructuring: per-rank partial multifrontal LU + interface LU on rank 0, "
                           "as preconditioner of distributed GMRES",
                 "max_front": info["maxfront"]}

    # end to end through host buffers (rank-local): H2D fields, assemble, D2H local values + rhs
    import ctypes as C
    vals_host = torch.empty(S.nnz * 2, dtype=torch.float64).pin_memory().numpy().view(np.complex128)
    rhs_host = torch.empty(S.ndofs * 2, dtype=torch.float64).pin_memory().numpy().view(np.complex128)

    inputs = prob.frequency_inputs(kw["lfactor"])       # host evaluation of the fields: outside the step

    def e2e_step():
        prob.apply_frequency_inputs(inputs)
        S.assemble()
        S._check(S.lib.vlfs_get_values(S.ctx, 0, vals_host.ctypes.data_as(C.c_void_p)))
        S._check(S.lib.vlfs_get_vector(S.ctx, 0, rhs_host.ctypes.data_as(C.c_void_p)))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    nrep = max(2, min(args.steps, 5))
    for _ in range(nrep):
        e2e_step()
    torch.cuda.synchronize()
    e_ms = allmax((time.perf_counter() - t0) * 1e3 / nrep)
    barrier()
    ncf = len(prob.dom_Gf.cells); nci = len(prob.dom_Gin.cells)
    e2e = {"value": nnz_global / (e_ms * 1e-3), "unit": "nnz/s", "ms_per_step": e_ms,
           "h2d_bytes_per_step": int(allsum((5 * ncf * 25 + 2 * nci * 25) * 8)),
           "d2h_bytes_per_step": int(allsum(S.nnz * sT + S.ndofs * sT))}
    if rank == 0:
        out = {
            "metric": "assembly nnz/s (Yago 3D; + SpMV GB/s + solve s/step in `spmv`/`solve`)",
            "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (complex128 matrix)", "data": "synthetic (deterministic Yago mesh, analytic coefficients)",
            "config": {"workload": "%s x%d along x: Yago 5-4-1 %s, N=%d, nnz=%d (global), rows in %d x-slabs; per-rank inputs > L2, no flush"
                       % (args.workload, world, json.dumps(kw), N_global, nnz_global, world),
                       "parallelism": "row-partitioned x-slabs, owner-computes assembly, NCCL halo exchange + all-reduce"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(l2 - l1), "roofline": roof,
            "kernels_ms": prof, "spmv": spmv, "solve": solve, "cpu_baseline": None, "host_setup_s": setup_s,
        }
        emit(out)
    dist.destroy_process_group()


def cpu_solve_baseline():
    """SpMV and sparse LU of the CPU path on a bounded sample (1/8 of Y1 in x: SuperLU needs ~40 s
    there and grows superlinearly; UMFPACK, what Gridap's LUSolver calls, is the same algorithm class)."""
    import numpy as np
    import scipy.sparse.linalg as spla
    from oracle.cases.yago import YagoCase
    kw = dict(nx=8, ny=2, nz=3, order=4, lfactor=0.4, dfactor=4)
    A, b = YagoCase(**kw).assemble()
    A = A.tocsr()
    n = A.shape[0]
    x = np.sin(0.1 * np.arange(n)) + 1j * np.cos(0.1 * np.arange(n))
    A @ x
    t0 = time.time()
    for _ in range(10):
        A @ x
    sp_s = (time.time() - t0) / 10
    sp_bytes = A.nnz * 20 + 4 * (n + 1) + 2 * n * 16
    t0 = time.time()
    lu = spla.splu(A.tocsc())
    xs = lu.solve(b)
    lu_s = time.time() - t0
    return {"sample": "Yago %s: N=%d nnz=%d" % (json.dumps(kw), n, A.nnz), "cores": 1,
            "spmv_GBs": sp_bytes / sp_s / 1e9, "spmv_ms": sp_s * 1e3,
            "solve_s_per_step": lu_s, "solver": "scipy.sparse.linalg.splu (SuperLU) + triangular solves",
            "rel_residual": float(np.linalg.norm(A @ xs - b) / np.linalg.norm(b))}


_cpu_problem_cache = {}


def cpu_baseline(sample=None, with_solve=False, threads=0):
    """The compiled CPU restatement of the assembly (oracle/c/asm_cpu.cpp: quadrature of every
    cell's element blocks from the term descriptors, COO fill in Gridap's order, counting sort +
    per-row sort + duplicate sums = `sparse(I,J,V)`) on all host threads, on the same workload as
    the GPU arm.  The tables and descriptors come from the host mirror with the recording back end
    (no GPU involved; tests/test_host_tables.py ties them to the oracle's literal weak forms) and
    are built once, outside the timed region - as they are for the GPU arm."""
    import numpy as np
    from oracle.c_assembly import CSystem
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    kw = sample or WORKLOADS["Y1"]
    key = json.dumps(kw, sort_keys=True)
    if key not in _cpu_problem_cache:
        _cpu_problem_cache.clear()
        _cpu_problem_cache[key] = YagoProblem(Yago_freq_domain_params(**kw), system_cls=CSystem)
    prob = _cpu_problem_cache[key]
    prob.sys.threads = threads
    t0 = time.time()
    prob.sys.assemble()
    wall = time.time() - t0
    ms = prob.sys.ms
    dt = ms["total"] * 1e-3
    nnz, n = prob.sys.nnz, prob.X.ndofs
    cores = threads if threads >= 1 else (os.cpu_count() or 1)
    spmv = None
    try:                                                     # threaded CSR SpMV on the matrix just assembled
        from oracle.c_assembly import spmv_ms
        A = prob.sys.mats[0]
        x = np.sin(0.1 * np.arange(n)) + 1j * np.cos(0.1 * np.arange(n))
        sp_ms, _ = spmv_ms(A, x, reps=5, threads=threads)
        sp_bytes = nnz * 20 + 4 * (n + 1) + 2 * n * 16
        spmv = {"ms": sp_ms, "GBs": sp_bytes / (sp_ms * 1e-3) / 1e9, "cores": cores}
    except Exception as e:                                   # never let a CPU leg break the bench line
        spmv = {"error": repr(e)}
    prob.sys.mats = prob.sys.vecs = None                     # 1.2 GB of results: not needed any more
    extra = {}
    if with_solve:
        try:
            extra = {"spmv_solve": cpu_solve_baseline()}
        except Exception as e:
            extra = {"spmv_solve": {"error": repr(e)}}
    return {**extra, "spmv": spmv, "value": nnz / dt, "unit": "nnz/s", "cores": cores, "kind": "port",
            "N": n, "nnz": nnz, "ncoo": prob.sys.ncoo,
            "note": "compiled restatement (C++, host threads) of the Gridap algorithm class; Gridap 0.17 itself "
                    "integrates and assembles in ONE Julia thread (SURVEY.md §3) - the NumPy literal-form oracle "
                    "on one core does 0.9 M nnz/s",
            "sample": "the whole workload: Yago %s, N=%d nnz=%d ncoo=%d in %.2f s (integration + COO fill %.2f s, "
                      "sparse(I,J,V) %.2f s; %.2f s wall with result copies)"
                      % (json.dumps(kw), n, nnz, prob.sys.ncoo, dt, ms["integrate_and_fill"] * 1e-3,
                         ms["compress"] * 1e-3, wall), "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kw = WORKLOADS[args.workload]
    vals = []
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_baseline(kw)
    c = None
    nst = max(1, min(args.steps, 10))
    for i in range(nst):
        c = cpu_baseline(kw, with_solve=(i == nst - 1 and not args.no_cpu_solve))
        vals.append(c["value"])
    v = sum(vals) / len(vals)
    c["value"] = v
    out = {"impl": "reference", "metric": "assembly nnz/s (Yago 3D; + SpMV GB/s + solve s/step in `spmv`/`solve`)",
           "value": v, "unit": "nnz/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": len(vals),
           "warmup": 1, "ms_per_step": c["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64 (complex128 matrix)", "data": "synthetic",
           "config": {"workload": "%s: Yago 5-4-1 %s, N=%d, nnz=%d, ncoo=%d; CPU restatement on %d host threads, %s"
                                  % (args.workload, json.dumps(kw), c["N"], c["nnz"], c["ncoo"], c["cores"], c["sample"]),
                      "parallelism": "host threads"},
           "cpu_baseline": c,
           "e2e": {"value": v, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def main():
    # Libraries (NCCL banner, torchrun notices) may write to fd 1: keep the real stdout for the
    # ONE JSON line and send everything else to stderr.
    global _real_stdout
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="Y1", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-solve", action="store_true", help="skip the CPU SpMV / sparse LU sample (~50 s)")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--dist-maxit", type=int, default=120, help="GMRES iterations of the N>1 solve")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_ours_distributed(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
