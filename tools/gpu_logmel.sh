#!/bin/bash
# log-mel front end: parity tests, timing alone (full clock), clock sensitivity, ncu counters of one launch
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "logmel or mel_filters" -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/logmel_tests.log
echo "== tests exit ${PIPESTATUS[0]}"; tail -5 gpurun_out/logmel_tests.log
for n in 1024 1136 128; do timeout 120 python tools/run_logmel.py $n 2>&1 | tail -1; done | tee gpurun_out/logmel_timing.txt
timeout 120 python tools/logmel_clock_check.py 2>&1 | tail -6 | tee -a gpurun_out/logmel_timing.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel -c 1 -s 6 -o gpurun_out/logmel_fused -f python tools/run_logmel.py 1024 > gpurun_out/logmel_ncu.log 2>&1
echo "ncu exit $?"
python tools/ncu_source_lines.py gpurun_out/logmel_fused.ncu-rep 20 | cut -c1-150
