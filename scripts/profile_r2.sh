#!/bin/bash
# Round-2 profile evidence (one B200): plain run first, then the launch list and the full captures of the
# same command.  Outputs under gpurun_out/; scripts/ncu_summary.py turns them into profiles/r2_*.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-solve"
$B > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gather_chunks|expand_rows' -c 8 -o gpurun_out/r2_prof_numeric_pass $B > gpurun_out/r2_ncu_numeric.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'csr_spmv' -c 6 -o gpurun_out/r2_prof_spmv $B > gpurun_out/r2_ncu_spmv.log 2>&1
M="python bench.py --workload M2 --steps 2 --warmup 3 --no-cpu --no-solve"
$M > gpurun_out/r2_plain_m2.json 2> gpurun_out/r2_plain_m2.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:'integrate_cells' -c 8 -o gpurun_out/r2_prof_integrate_m2 $M > gpurun_out/r2_ncu_integrate.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_m2.csv $M > gpurun_out/r2_ncu_launches_m2.log 2>&1
ls -la gpurun_out/r2_*
