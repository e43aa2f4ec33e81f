#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 hot path (assembly nnz/s + SpMV GB/s + solve
s/step on the Yago 3-D case, BASELINE.json) with the CPU restatement timed beside it.

  python bench.py --gpus N --steps K --warmup W [--workload Y1] [--scaling weak|strong]   (our arm)
  python bench.py --impl reference --gpus N --steps K --warmup W                          (CPU reference arm)

A "step" is one numeric pass of `AffineFEOperator(a,b,X,Y)` for one frequency: every CSR value of
the monolithic matrix written (zero fill included) with all term contributions summed + the
right-hand side.  `value` = nnz / step time, inputs resident in HBM.  `e2e` is the same metric
through the host-buffer C-ABI path with the frequency changing every step: coefficient fields
copied host->device, the assembly, and the matrix values + rhs copied device->host, all inside the
timed region.  SpMV (complex and real) and the solve are timed in the same run (`spmv`, `spmv_real`,
`solve`).  N > 1: one process per GPU (torchrun); the library owns the NCCL communicator, the halo
exchange, the Krylov all-reduce and the interface chain (csrc/dist.cu) - torch.distributed only
carries the 128-byte NCCL id, the barriers and the max-over-ranks of the timings.
Workloads M1/M2/M3 (MultiGeo 5-5-1 on the gmsh mesh, refined 0/1/2 times) exercise the general
quadrature kernel `integrate_cells`, reported against the MEASURED FP64 FMA rate of the device.
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
    # Yago 5-4-1 parameters   [scripts/5-4-1-Yago.jl:17-46; SURVEY.md §8(d)]
    "Y0": dict(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3),
    "Ymid": dict(nx=8, ny=2, nz=2, order=4, lfactor=0.4, dfactor=4),
    "Y1": dict(nx=32, ny=4, nz=3, order=4, lfactor=0.4, dfactor=4),
    "Y2": dict(nx=64, ny=8, nz=6, order=4, lfactor=0.4, dfactor=4),
    # MultiGeo 5-5-1 on the fine gmsh mesh, order 4, `refine` levels of uniform refinement  [scripts/5-5-1-MultiGeo.jl:28-37]
    "M1": dict(multigeo=True, order=4, refine=0),
    "M2": dict(multigeo=True, order=4, refine=1),
    "M3": dict(multigeo=True, order=4, refine=2),
}
METRIC = "assembly nnz/s (Yago 3D; + SpMV GB/s + solve s/step in `spmv`/`solve`)"

_real_stdout = None


def emit(obj):
    f = _real_stdout or sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def workload_label(name, scale=1, scaling="weak"):
    """The same string in both arms (the driver compares `config`); arm-specific notes go elsewhere."""
    kw = WORKLOADS[name]
    if kw.get("multigeo"):
        return "%s: MultiGeo 5-5-1 models/multi_geo.msh order %d refined %d x" % (name, kw["order"], kw["refine"])
    if scale > 1 and scaling == "weak":
        kw = dict(kw, nx=kw["nx"] * scale)
        return "%s x%d along x: Yago 5-4-1 %s" % (name, scale, json.dumps(kw, sort_keys=True))
    return "%s: Yago 5-4-1 %s" % (name, json.dumps(kw, sort_keys=True))


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel from the committed
    `ncu --set full` capture (profiles/r2_traffic.json, else r1), or None."""
    for n in ("r2_traffic.json", "r1_traffic.json"):
        try:
            v = json.load(open(os.path.join(ROOT, "profiles", n))).get(kernel)
            if v is not None:
                return v
        except Exception:
            pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def bind_to_gpu_numa_node(index):
    """Pinned host buffers should live on the NUMA node of the rank's GPU (the device->host copy of
    the matrix values is the e2e step).  Best effort: silently keeps the inherited affinity."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


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


# ---------------------------------------------------------------------------------------------
# problem construction and the byte / flop counts of DESIGN.md §4
# ---------------------------------------------------------------------------------------------
def build_problem(name, device=0, system_cls=None, scale=1, scaling="weak"):
    kw = dict(WORKLOADS[name])
    extra = {} if system_cls is None else {"system_cls": system_cls}
    if kw.pop("multigeo", False):
        from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
        mesh = os.path.join(ROOT, "tests", "golden", "multi_geo.msh")
        return MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=mesh, order=kw["order"], refine=kw["refine"]),
                               device=device, **extra), kw
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    if scale > 1 and scaling == "weak":
        kw["nx"] = kw["nx"] * scale
    return YagoProblem(Yago_freq_domain_params(**kw), device=device, **extra), kw


def domain_tables_bytes(S):
    """DOF maps + coordinates + coefficient fields read once per pass: Σ ncells (4 ndofs + 8 D nv + 8 nq nw)."""
    return sum(d.table_bytes for d in getattr(S, "domains", []) if hasattr(d, "table_bytes"))


def pass_bytes(nnz, ncoo, sT):
    """every CSR value written once + one 4-byte record per COO entry read once; the
    SAME formula for the kernel's roofline at N = 1 and N > 1."""
    return nnz * sT + 4 * ncoo


def quadrature_flops(S):
    """FP64 flops of the general-quadrature kernel per pass: Σ_domains ncells nq Σ_terms 2 ncomp nr nc
    (x2 when the coefficient has a real and an imaginary part; the register-tiled contraction only -
    operator-table evaluation and geometry are not counted)."""
    return sum(d.ncells * d.nq * d.term_flops for d in getattr(S, "domains", []) if hasattr(d, "term_flops"))


def numeric_pass(prof):
    """(kernel name, ms) of the numeric pass: all launches of the pass together (template gather in
    two phases + row expansion, or the plain chunk gather / the slot-centric fallback)."""
    names = sorted({n.split(":")[0] for n, t in prof if n.startswith(("gather_chunks", "expand_rows")) and t > 0})
    if names:
        return "+".join(names), sum(t for n, t in prof if n.split(":")[0] in names)
    for n, t in prof:
        if n.startswith("gather_reference"):
            return "gather_reference", t
    return max(prof, key=lambda kv: kv[1])


def run_ours(args):
    import numpy as np
    import torch
    import vlfs_b200
    from vlfs_b200 import _lib as L
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local)
    t0 = time.time()
    prob, kw = build_problem(args.workload, device=local)
    setup_s = time.time() - t0
    multigeo = bool(WORKLOADS[args.workload].get("multigeo"))
    S = prob.sys
    nnz, ncoo, N = S.nnz, S.ncoo, S.ndofs
    sT = 16
    peak, peak_kind = measured_peaks()

    def sync():
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        S.assemble()
    clocks = ClockSampler(local)
    sync()
    clocks.start()
    launches1 = S.launches()
    ms = S.assemble_device_ms(reps=args.steps)
    launches2 = S.launches()
    sync()
    clk = clocks.stop()
    value = nnz / (ms * 1e-3)

    # per-kernel times (CUDA events on the launching stream) for the roofline of the dominant kernel
    prof = S.assemble_profile(reps=3)
    prof_total = sum(t for _, t in prof)
    fp64_peak = S.fp64_peak_tflops()
    dom_name, dom_ms = max(prof, key=lambda kv: kv[1])
    if dom_name.startswith("integrate_cells"):
        # general quadrature on every cell (unstructured mesh): FP64-FMA bound (SURVEY.md §7 hard part 4)
        integ_ms = sum(t for n, t in prof if n.startswith("integrate_cells"))
        flops = quadrature_flops(S)
        roof = {"bound": "fp64", "kernel": "integrate_cells", "achieved": flops / (integ_ms * 1e-3) / 1e12,
                "peak": fp64_peak, "peak_kind": "measured in this run (vlfs_measure_fp64_peak: dependent-free DFMA chains)",
                "unit": "TFLOP/s", "traffic": None, "share_of_step": integ_ms / prof_total, "algorithmic_flops": flops}
    else:
        dom_name, dom_ms = numeric_pass(prof)
        kbytes = pass_bytes(nnz, ncoo, sT)
        roof = {"bound": "hbm", "kernel": dom_name, "achieved": kbytes / (dom_ms * 1e-3) / 1e9, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s",
                "traffic": ncu_traffic(dom_name) if args.workload == "Y1" else None,   # the ncu capture is of Y1
                "share_of_step": dom_ms / prof_total, "algorithmic_bytes": kbytes,
                "formula": "nnz*sT + 4*ncoo (SURVEY.md §8d); both launches of a two-phase kernel together"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    step_bytes = pass_bytes(nnz, ncoo, sT) + domain_tables_bytes(S)

    # SpMV (device resident, x fixed = (sin 0.1 i, cos 0.1 i))
    S.spmv_device_ms(reps=3)
    sp_ms = S.spmv_device_ms(reps=max(args.steps, 10))
    sp_bytes = nnz * (sT + 4) + 4 * (N + 1) + 2 * N * sT
    spmv = {"ms": sp_ms, "GBs": sp_bytes / (sp_ms * 1e-3) / 1e9, "bytes": sp_bytes,
            "traffic": ncu_traffic("csr_spmv") if args.workload == "Y1" else None}
    spmv["frac"] = spmv["GBs"] / peak

    # real CSR SpMV (north star 3, the Newmark products) at HBM scale: the real part of the same matrix
    spmv_real = None
    if not args.no_real_spmv and nnz * 8 < 40e9:
        try:
            A = S.get_matrix(0)
            A.data = np.ascontiguousarray(A.data.real)
            Sr = vlfs_b200.B200System.from_csr(A, device=local)
            del A
            Sr.spmv_device_ms(reps=3)
            r_ms = Sr.spmv_device_ms(reps=max(args.steps, 10))
            r_bytes = nnz * 12 + 4 * (N + 1) + 2 * N * 8
            spmv_real = {"ms": r_ms, "GBs": r_bytes / (r_ms * 1e-3) / 1e9, "bytes": r_bytes,
                         "frac": r_bytes / (r_ms * 1e-3) / 1e9 / peak,
                         "matrix": "real part of the workload's matrix (same pattern, imported with vlfs_import_csr)"}
            Sr.close()
        except Exception as e:
            spmv_real = {"error": repr(e)}

    # solve: sparse LU factorisation + GMRES refinement per frequency
    solve = None
    if not args.no_solve:
        try:
            t0 = time.time()
            S.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=20, maxit=40, rtol=1e-10)
            first_s = time.time() - t0
            samples = []
            for _ in range(3):                      # one frequency step = refactorise + solve; median of 3
                t0 = time.time()
                S.refactor(0)
                x, st = S.solve(vec=0)
                samples.append(time.time() - t0)
            info = S.direct_info()
            pert, bad = S.direct_status()
            tf = info["flops"] / (info["numeric_ms"] * 1e-3) / 1e12
            solve = {"s_per_step": sorted(samples)[1], "s_per_step_samples": samples, "first_setup_s": first_s,
                     "symbolic_ms": info["symbolic_ms"], "numeric_ms": info["numeric_ms"],
                     "gmres_iterations": st["iterations"], "rel_residual": st["rel_residual"], "converged": st["converged"],
                     "backward_error": st.get("backward_error"), "at_attainable_accuracy": st.get("at_attainable_accuracy"),
                     "solve_ms": st["solve_ms"], "factor_TFLOPs": tf, "fp64_peak_TFLOPs_measured": fp64_peak,
                     "factor_frac_of_fp64_peak": tf / fp64_peak, "pivots_perturbed": pert, "pivots_nonfinite": bad,
                     "fill_entries": info["fill"], "fronts": info["nfronts"], "max_front": info["maxfront"]}
            if not multigeo:
                xs, rao = prob.rao(x)
                solve["rao_max"] = float(rao.max()) if len(rao) else None
            # solution error on a manufactured problem: b := A x_true, solve, compare (what double precision
            # can deliver on this matrix is cond(A) * backward error; the 1e-8 gate of the parity tests is
            # taken against an extended-precision refinement of the oracle's LU solution)
            xt = (np.cos(0.001 * np.arange(N)) + 1j * np.sin(0.003 * np.arange(N))) * (abs(x).mean() or 1.0)
            bt = S.spmv(xt)
            xm, stm = S.solve(b=bt)
            solve["manufactured_solution_rel_error"] = float(np.linalg.norm(xm - xt) / np.linalg.norm(xt))
            solve["manufactured_rel_residual"] = stm["rel_residual"]
        except Exception as e:
            solve = {"error": repr(e)}

    # end to end through host buffers, a NEW frequency every step: H2D of the ω-dependent coefficient
    # fields and scalars, assemble (re-classification included when a field a bilinear term uses changes),
    # D2H of the matrix values + rhs
    import ctypes as C
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    vals_host = torch.empty(nnz * 2, dtype=torch.float64).pin_memory().numpy().view(np.complex128)
    rhs_host = torch.empty(N * 2, dtype=torch.float64).pin_memory().numpy().view(np.complex128)
    parts = [0.0, 0.0, 0.0]
    if multigeo:
        inputs, h2d = [None], 0
    else:
        inputs = []
        for lf in (0.4, 0.6):                 # the sweep of scripts/5-4-1-Yago.jl:33-71, host evaluation outside the step
            inp = prob.frequency_inputs(lf)
            for key in ("weights_Gf", "weights_Gin"):
                inp[key] = [pin(a) for a in inp[key]]
            inputs.append(inp)
        h2d = (5 * prob.Gf.ncells * 25 + 2 * prob.Gin.ncells * 25) * 8

    def e2e_step(k):
        t_a = time.perf_counter()
        if inputs[0] is not None:
            prob.apply_frequency_inputs(inputs[k % len(inputs)])
        t_b = time.perf_counter()
        S.assemble()
        t_c = time.perf_counter()
        S._check(S.lib.vlfs_get_values(S.ctx, 0, vals_host.ctypes.data_as(C.c_void_p)))
        S._check(S.lib.vlfs_get_vector(S.ctx, 0, rhs_host.ctypes.data_as(C.c_void_p)))
        t_d = time.perf_counter()
        parts[0] += t_b - t_a; parts[1] += t_c - t_b; parts[2] += t_d - t_c
    d2h = nnz * sT + N * sT
    e2e_step(0); e2e_step(1)
    sync()
    nrep = max(2, min(args.steps, 6))
    parts[:] = [0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    for k in range(nrep):
        e2e_step(k)
    sync()
    e_ms = (time.perf_counter() - t0) * 1e3 / nrep
    e2e = {"value": nnz / (e_ms * 1e-3), "unit": "nnz/s", "ms_per_step": e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "frequency_changes_every_step": inputs[0] is not None,
           "parts_ms": {"set_coefficients_h2d": parts[0] * 1e3 / nrep, "assemble": parts[1] * 1e3 / nrep,
                        "values_d2h": parts[2] * 1e3 / nrep}}

    cpu = None
    if not args.no_cpu and not multigeo:
        try:
            cpu = cpu_baseline(WORKLOADS[args.workload], with_solve=not args.no_cpu_solve)
        except Exception as e:              # the CPU leg must not take the GPU numbers down with it
            cpu = {"error": repr(e), "kind": "port"}

    out = {
        "metric": METRIC, "value": value, "unit": "nnz/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (complex128 matrix)", "data": "synthetic (deterministic mesh, analytic coefficients)",
        "config": {"workload": workload_label(args.workload), "parallelism": "single"},
        "sizes": {"N": N, "nnz": nnz, "ncoo": ncoo, "inputs_vs_L2": "values %.2f GB + records %.2f GB > 126 MB L2, no flush"
                  % (nnz * sT / 1e9, ncoo * 4 / 1e9)},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches2 - launches1),
        "roofline": roof, "step_GBs": step_bytes / (ms * 1e-3) / 1e9, "step_frac": step_bytes / (ms * 1e-3) / 1e9 / peak,
        "kernels_ms": prof, "pattern_build_ms": getattr(S, "pattern_build_ms", None),
        "spmv": spmv, "spmv_real": spmv_real, "solve": solve, "cpu_baseline": cpu,
        "host_setup_s": setup_s, "numa_node": numa,
    }
    emit(out)


def run_ours_distributed(args):
    """N > 1: one process per GPU.  weak: the Yago mesh refined N x along x (per-GPU work = Y1);
    strong: the SAME mesh split into N x-slabs of cells.  The cells are bisected by centroid, a DOF
    belongs to the highest part that touches it, every rank assembles its own rows; SpMV, solve and
    all communication run inside libvlfs_b200 (csrc/dist.cu)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import vlfs_b200
    from vlfs_b200 import _lib as L, dist as D
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    multigeo = bool(WORKLOADS[args.workload].get("multigeo"))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    import datetime
    # (a rank that fails must not keep the others in a collective for the default 10 minutes)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    uid = [D.MultiSystem.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sT = 16
    peak, peak_kind = measured_peaks()
    t0 = time.time()
    mk = lambda *a, **k: D.MultiSystem(*a, nranks=world, local_ranks=[rank], devices=[local], unique_id=uid[0],
                                       partition="cells", rcb_axes=(0,), **{kk: v for kk, v in k.items() if kk != "device"})
    prob, kw = build_problem(args.workload, device=local, system_cls=mk, scale=world, scaling=args.scaling)
    setup_s = time.time() - t0
    M = prob.sys
    part = M.parts[0]
    S = part.local

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

    nnz_global = int(allsum(M.nnz_owned[0]))
    N_global = M.ndofs
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
    dom_name, dom_ms = numeric_pass(prof)
    kbytes = pass_bytes(S.nnz, S.ncoo, sT)                      # this rank's local matrix (ghost rows included)
    roof = {"bound": "hbm", "kernel": dom_name + " (rank 0)", "achieved": kbytes / (dom_ms * 1e-3) / 1e9,
            "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "traffic": None, "algorithmic_bytes": kbytes,
            "share_of_step": dom_ms / sum(t for _, t in prof),
            "formula": "nnz*sT + 4*ncoo of the rank's local system (SURVEY.md §8d), as at N = 1"}
    roof["frac"] = roof["achieved"] / peak

    # distributed SpMV inside the library: halo (pack + ncclSend/Recv on a side stream) overlapped with the
    # rows without ghost columns, CUDA events on the rank's stream, max over ranks
    M.spmv_device_ms(reps=3)
    barrier()
    sp_ms = allmax(M.spmv_device_ms(reps=max(args.steps, 10)))
    barrier()
    sp_bytes = nnz_global * (sT + 4) + 4 * (N_global + 1) + 2 * N_global * sT
    info0 = M.dist_info()[0]
    spmv = {"ms": sp_ms, "GBs": sp_bytes / (sp_ms * 1e-3) / 1e9, "bytes": sp_bytes,
            "frac": sp_bytes / (sp_ms * 1e-3) / 1e9 / world / peak,
            "halo_values_sent_rank0": int(part.halo[1][-1]), "rows_with_ghost_columns_rank0": info0["boundary_rows"]}

    solve = None
    if not args.no_solve:
        try:
            barrier()
            t0 = time.time()
            M.solver_setup(0, restart=20, maxit=40, rtol=1e-10)
            torch.cuda.synchronize()
            first_s = allmax(time.time() - t0)
            samples = []
            for _ in range(3):
                barrier()
                t0 = time.time()
                M.refactor(0)
                x, st = M.solve(vec=0)
                torch.cuda.synchronize()
                samples.append(allmax(time.time() - t0))
            info = S.direct_info()
            di = M.dist_info()[0]
            solve = {"s_per_step": sorted(samples)[1], "s_per_step_samples": samples, "first_setup_s": first_s,
                     "symbolic_ms": info["symbolic_ms"], "numeric_ms_rank0": info["numeric_ms"],
                     "interface_chain_ms_rank0": di["chain_ms"], "cut_unknowns_rank0": [di["n_left"], di["n_right"]],
                     "gmres_iterations": st["iterations"], "rel_residual": st["rel_residual"], "converged": st["converged"],
                     "backward_error": st.get("backward_error"), "at_attainable_accuracy": st.get("at_attainable_accuracy"),
                     "solve_ms": st["solve_ms"], "max_front": info["maxfront"],
                     "method": "substructuring on the chain of slab cuts inside libvlfs_b200: local multifrontal LU with "
                               "the cuts kept, every rank factorises its own link of the interface chain; GMRES with one "
                               "ncclAllReduce per iteration"}
            # the same mesh solved on ONE GPU (rank 0; its LU fits beside the rank's own data up to N = 2): solution error of the distributed solve
            if world <= 2 and not args.no_single_check and not multigeo:
                xg = torch.from_numpy(np.ascontiguousarray(x).view(np.float64)).cuda()
                dist.all_reduce(xg)                 # owned entries are disjoint, the others zero
                if rank == 0:
                    try:                            # (an exception here must not leave the other ranks in the barrier)
                        xd = xg.cpu().numpy().view(np.complex128)
                        p1, _ = build_problem(args.workload, device=local, scale=world, scaling=args.scaling)
                        p1.assemble()
                        p1.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=20, maxit=40, rtol=1e-10)
                        x1, st1 = p1.sys.solve(vec=0)
                        solve["vs_single_gpu_solve_of_the_same_mesh"] = {
                            "solution_rel_error": float(np.linalg.norm(xd - x1) / np.linalg.norm(x1)),
                            "single_gpu_rel_residual": st1["rel_residual"], "single_gpu_backward_error": st1.get("backward_error")}
                        p1.sys.close()
                        del p1
                    except Exception as e:
                        solve["vs_single_gpu_solve_of_the_same_mesh"] = {"error": repr(e)}
                del xg
                barrier()
        except Exception as e:
            solve = {"error": repr(e)}

    # end to end through host buffers (rank-local): H2D fields, assemble, D2H local values + rhs
    import ctypes as C
    vals_host = torch.empty(S.nnz * 2, dtype=torch.float64).pin_memory().numpy().view(np.complex128)
    rhs_host = torch.empty(S.ndofs * 2, dtype=torch.float64).pin_memory().numpy().view(np.complex128)
    # host evaluation of the fields: outside the step (MultiGeo: one frequency, nothing to upload)
    inputs = [None] if multigeo else [prob.frequency_inputs(lf) for lf in (0.4, 0.6)]

    def e2e_step(k):
        if inputs[0] is not None:
            prob.apply_frequency_inputs(inputs[k % 2])
        S.assemble()
        S._check(S.lib.vlfs_get_values(S.ctx, 0, vals_host.ctypes.data_as(C.c_void_p)))
        S._check(S.lib.vlfs_get_vector(S.ctx, 0, rhs_host.ctypes.data_as(C.c_void_p)))
    e2e_step(0); e2e_step(1)
    barrier()
    t0 = time.perf_counter()
    nrep = max(2, min(args.steps, 6))
    for k in range(nrep):
        e2e_step(k)
    torch.cuda.synchronize()
    e_ms = allmax((time.perf_counter() - t0) * 1e3 / nrep)
    barrier()
    h2d_rank = 0 if multigeo else (5 * len(prob.dom_Gf.cells) * 25 + 2 * len(prob.dom_Gin.cells) * 25) * 8
    e2e = {"value": nnz_global / (e_ms * 1e-3), "unit": "nnz/s", "ms_per_step": e_ms,
           "h2d_bytes_per_step": int(allsum(h2d_rank)),
           "d2h_bytes_per_step": int(allsum(S.nnz * sT + S.ndofs * sT)), "frequency_changes_every_step": not multigeo}
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": ("strong" if multigeo else args.scaling), "vs_baseline": None,
            "dtype": "f64 (complex128 matrix)", "data": "synthetic (deterministic mesh, analytic coefficients)",
            "config": {"workload": workload_label(args.workload, world, args.scaling),
                       "parallelism": "cells bisected into %d x-slabs, owner-computes rows; halo exchange, all-reduce and "
                                      "interface chain inside libvlfs_b200 over NCCL" % world},
            "sizes": {"N": N_global, "nnz": nnz_global, "nnz_local_rank0": S.nnz, "ncoo_local_rank0": S.ncoo},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(l2 - l1), "roofline": roof,
            "kernels_ms": prof, "pattern_build_ms": getattr(S, "pattern_build_ms", None), "spmv": spmv, "solve": solve,
            "cpu_baseline": None, "host_setup_s": setup_s, "numa_node": numa,
        }
        emit(out)
    M.close()
    dist.destroy_process_group()


def cpu_solve_baseline():
    """SpMV and sparse LU of the CPU path on a bounded sample (1/8 of Y1 in x: SuperLU needs ~40 s
    there and grows superlinearly; UMFPACK, what Gridap's LUSolver calls, is the same algorithm class).
    The full Y1 matrix (N = 1.02 M, 77.5 M complex entries) does not factorise on the host within the
    bench's time budget: SuperLU's fill on this sample is already 216 M entries at 1/8 of the size and
    grows ~ n log n on the quasi-2-D mesh, i.e. > 2 G entries (> 30 GB) and tens of minutes."""
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
    return {"sample": "Yago %s: N=%d nnz=%d (1/8 of Y1; the full matrix does not factorise on the host in the bench's budget)"
                      % (json.dumps(kw), n, A.nnz), "cores": 1,
            "spmv_GBs": sp_bytes / sp_s / 1e9, "spmv_ms": sp_s * 1e3,
            "solve_s_per_step": lu_s, "solver": "scipy.sparse.linalg.splu (SuperLU) + triangular solves",
            "fill_entries": int(lu.L.nnz + lu.U.nnz),
            # same-config estimate: the sample is Y1 coarsened 4 x 2 in the plane (quasi-2-D mesh, 3 layers deep);
            # a sparse LU of a 2-D-like mesh costs ~ n^1.5 flops (nested dissection bound; SuperLU's COLAMD
            # ordering does not do better), so the whole workload is >= 8^1.5 = 22.6 x this time on one core
            "extrapolated_full_workload_s": lu_s * 8 ** 1.5,
            "extrapolation_basis": "measured sample time x (N_Y1/N_sample)^1.5 = x 22.6 (flops of a sparse LU on a "
                                   "quasi-2-D mesh); lower bound, SuperLU is single-threaded here",
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
    name = args.workload if not WORKLOADS[args.workload].get("multigeo") else "Y1"
    kw = WORKLOADS[name]
    for _ in range(max(0, args.warmup)):
        cpu_baseline(kw)
    secs = []
    c = None
    nst = max(1, args.steps)
    for i in range(nst):
        c = cpu_baseline(kw, with_solve=(i == nst - 1 and not args.no_cpu_solve))
        secs.append(c["seconds"])
    v = c["nnz"] * len(secs) / sum(secs)                   # units processed / time of exactly K steps
    c["value"] = v
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "nnz/s", "n_gpus": world, "steps": len(secs),
           "warmup": max(0, args.warmup), "ms_per_step": sum(secs) / len(secs) * 1e3, "higher_is_better": True,
           "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f64 (complex128 matrix)",
           "data": "synthetic (deterministic mesh, analytic coefficients)",
           "config": {"workload": workload_label(name, world, args.scaling), "parallelism": "single" if world == 1 else
                      "cells bisected into %d x-slabs, owner-computes rows; halo exchange, all-reduce and "
                      "interface chain inside libvlfs_b200 over NCCL" % world},
           "arm_notes": "CPU restatement on %d host threads: %s (at N > 1 the CPU arm still processes the N = 1 mesh)"
                        % (c["cores"], c["sample"]),
           "sizes": {"N": c["N"], "nnz": c["nnz"], "ncoo": c["ncoo"]},
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = mesh refined N x along x (per-GPU work fixed); strong = the same mesh split N ways")
    ap.add_argument("--no-cpu-solve", action="store_true", help="skip the CPU SpMV / sparse LU sample (~50 s)")
    ap.add_argument("--no-single-check", action="store_true", help="N > 1: skip the single-GPU solve of the same mesh")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-real-spmv", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_ours_distributed(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
