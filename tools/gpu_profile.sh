#!/bin/bash
# launch list (per-launch device time) and full captures of the two dominant kernels for a short bench run
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --hours 0.15 --no-cpu-baseline --no-extra-workloads --no-side-kernels"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 2 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc5 -s 1 -c 1 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
tail -c 600 gpurun_out/plain.log
ls -la gpurun_out
