#!/bin/bash
# sweep of the numeric pass's variants on the Y1 workload (one process per setting)
out=gpurun_out/asm_sweep.jsonl; : > $out
python -m pytest tests/test_gpu_numeric_passes.py tests/test_gpu_yago.py tests/test_gpu_cases.py -m gpu -x -q 2>&1 | tail -15
for cfg in "VLFS_DEBUG=1" "VLFS_ASM_TEMPLATES=0" "VLFS_ASM_CHUNKS=0"; do
  env $cfg python bench.py --steps 20 --warmup 3 --no-cpu --no-solve 2>gpurun_out/sweep.err | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'cfg':'$cfg','ms':j['ms_per_step'],'kernels':j['kernels_ms'],'frac':j['roofline']['frac'],'pattern_build_ms':j.get('pattern_build_ms')}))" >> $out
  grep "vlfs" gpurun_out/sweep.err | head -20
done
cat $out
