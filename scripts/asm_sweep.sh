#!/bin/bash
# sweep of the numeric pass's variants on the Y1 workload (one process per setting)
out=gpurun_out/asm_sweep.jsonl; : > $out
python -m pytest tests/test_gpu_yago.py tests/test_gpu_cases.py -m gpu -x -q 2>&1 | tail -3
for cfg in "VLFS_TILES_G=8" "VLFS_TILES_G=4" "VLFS_TILES_G=2"; do
  env $cfg python bench.py --steps 20 --warmup 3 --no-cpu --no-solve 2>gpurun_out/sweep.err | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'cfg':'$cfg','ms':j['ms_per_step'],'kernels':j['kernels_ms'],'frac':j['roofline']['frac']}))" >> $out
done
cat $out
