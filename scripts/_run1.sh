python -m pytest tests/test_gpu_numeric_passes.py -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 --no-cpu --no-solve > gpurun_out/b1.json 2> gpurun_out/b1.err
python -c "
import json
j=json.loads(open('gpurun_out/b1.json').read().strip().splitlines()[-1])
print(j['ms_per_step'], j['roofline']['frac'], j['kernels_ms'])"
ncu --set full --clock-control none --import-source on -k regex:expand_rows -c 2 -o gpurun_out/prof_expand python bench.py --steps 2 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu_e.log 2>&1
tail -2 gpurun_out/ncu_e.log | cut -c1-200
