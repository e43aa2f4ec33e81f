"""Turns `ncu -i X.ncu-rep --page raw --csv` into the two files under profiles/ that bench.py and
the notes cite: a per-launch summary of selected metrics and the DRAM bytes per launch.

usage: ncu -i gpurun_out/r1_prof_assembly_spmv.ncu-rep --page raw --csv > /tmp/raw.csv
       python scripts/ncu_summary.py /tmp/raw.csv profiles/r1_ncu_full_summary.csv profiles/r1_traffic.json
"""
import csv
import json
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
    for key, lab in (("build_class_tables", "class_tables:build"), ("integrate_cells", "class_tables:integrate"),
                     ("gather_compact", "gather_reference"), ("gather_reference", "gather_reference"), ("rhs_affine", "rhs_affine"), ("csr_spmv", "csr_spmv")):
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
    out = {"gather_reference": per.get("gather_reference"), "csr_spmv": per.get("csr_spmv"),
           "captured_launches": {lab: len(v) for lab, v in traffic.items()}, "per_launch_mean_bytes": per}
    json.dump(out, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:4])
