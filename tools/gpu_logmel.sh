#!/bin/bash
# log-mel front end: parity tests, timing of the fused kernel against the two-kernel form, ncu counters
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "logmel or mel_filters" -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/logmel_tests.log
echo "== tests exit ${PIPESTATUS[0]}"; tail -25 gpurun_out/logmel_tests.log
for n in 1024 1136 128; do
  echo "fused $n:"; timeout 120 python tools/run_logmel.py $n 2>&1 | tail -2
  echo "split $n:"; SEGMA_LOGMEL_SPLIT=1 timeout 120 python tools/run_logmel.py $n 2>&1 | tail -2
done | tee gpurun_out/logmel_timing.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel -c 2 -s 6 -o gpurun_out/logmel_fused -f python tools/run_logmel.py 1024 > gpurun_out/logmel_ncu.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/logmel_fused.ncu-rep --page raw --csv > gpurun_out/logmel_fused_raw.csv 2>/dev/null
