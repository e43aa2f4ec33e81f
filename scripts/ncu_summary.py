"""Turns `ncu -i X.ncu-rep --page raw --csv` into the two files under profiles/ that bench.py and
the notes cite: a per-launch summary of selected metrics and the DRAM bytes per launch.

usage: ncu -i gpurun_out/r1_prof_assembly_spmv.ncu-rep --page raw --csv > /tmp/raw.csv
       python scripts/ncu_summary.py /tmp/raw.csv profiles/r1_ncu_full_summary.csv profiles/r1_traffic.json
"""
import csv
import json
import os
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def label(name, seen):
    for key, lab in (("build_class_tables", "class_tables:build"), ("integrate_cells", "integrate_cells"),
                     ("gather_chunks", "gather_chunks"), ("expand_rows", "expand_rows"),
                     ("gather_compact", "gather_reference"), ("gather_reference", "gather_reference"), ("rhs_affine", "rhs_affine"),
                     ("csr_spmv_cplx", "csr_spmv"), ("csr_spmv_real", "csr_spmv_real"), ("schur_update", "schur_update")):
        if key in name:
            seen[lab] = seen.get(lab, 0) + 1
            return lab, seen[lab]
    return name[:40], 0


def main(raw, out_csv, out_json):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    seen, traffic, count = {}, {}, {}
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["label", "Kernel Name"] + [m for m, _ in cols])
        w.writerow(["unit", ""] + [units[i] for _, i in cols])
        for r in data:
            lab, k = label(r[kn], seen)
            w.writerow(["%s#%d" % (lab, k), r[kn]] + [r[i] for _, i in cols])
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            b = float(r[rd]) * SCALE[units[rd]] + float(r[wr]) * SCALE[units[wr]]
            traffic.setdefault(lab, []).append(b)
    # per launch: mean over the captured launches of the same kernel
    per = {lab: sum(v) / len(v) for lab, v in traffic.items()}
    out = {"captured_launches": {lab: len(v) for lab, v in traffic.items()}, "per_launch_mean_bytes": per}
    for k in ("gather_reference", "csr_spmv", "csr_spmv_real", "integrate_cells"):
        if k in per:
            out[k] = per[k]
    # the numeric pass of one assembly = 2 gather_chunks + 2 expand_rows launches (early / late): captured in
    # step order, so the first four launches are one pass
    if "gather_chunks" in traffic and "expand_rows" in traffic and len(traffic["gather_chunks"]) >= 2 and len(traffic["expand_rows"]) >= 2:
        out["expand_rows+gather_chunks"] = sum(traffic["gather_chunks"][:2]) + sum(traffic["expand_rows"][:2])
    if os.path.exists(out_json) and len(sys.argv) > 4 and sys.argv[4] == "merge":
        old = json.load(open(out_json))
        for k in ("captured_launches", "per_launch_mean_bytes"):
            old.setdefault(k, {}).update(out.pop(k))
        old.update(out)
        out = old
    json.dump(out, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:4])
