#!/bin/bash
# copies the bench lines of the last GPU runs (gpurun_out/, scratch) into profiles/ (tracked)
cd "$(dirname "$0")/.."
last() { f="gpurun_out/$1"; [ -s "$f" ] && tail -n 1 "$f" > "profiles/$2" && echo "$2"; }
last bench_n1.json r2_bench_n1.json
last bench_ref.json r2_bench_reference.json
last bench_n2.json r2_bench_n2.json
last bench_n4.json r2_bench_n4.json
last bench_n8.json r2_bench_n8.json
last bench_m2.json r2_bench_m2.json
last bench_m2_n2.json r2_bench_m2_n2.json
last bench_m3_n8.json r2_bench_m3_n8.json
last bench_y2_n1.json r2_bench_y2_n1.json
last bench_y2s_n2.json r2_bench_y2_strong_n2.json
last bench_y2s_n4.json r2_bench_y2_strong_n4.json
last bench_y2s_n8.json r2_bench_y2_strong_n8.json
[ -s gpurun_out/configs_r2.jsonl ] && cp gpurun_out/configs_r2.jsonl profiles/r2_configs_paper_resolution.jsonl
true
