#!/bin/bash
# sweep of the row-centric pass's tuning knobs on the Y1 workload (one process per setting)
out=gpurun_out/asm_sweep.jsonl; : > $out
python -m pytest tests/test_gpu_yago.py tests/test_gpu_cases.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for cfg in "VLFS_DEBUG=1" "VLFS_ROWS_CTAS=3"; do
  env $cfg python bench.py --steps 20 --warmup 3 --no-cpu --no-solve 2>gpurun_out/sweep.err | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'cfg':'$cfg','ms':j['ms_per_step'],'kernels':j['kernels_ms'],'frac':j['roofline']['frac']}))" >> $out
  grep "assemble_rows" gpurun_out/sweep.err | head -2
done
cat $out
