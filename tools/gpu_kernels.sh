#!/bin/bash
# run the per-kernel GPU parity tests in separate processes so one hang cannot hide the others
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tee gpurun_out/gpu.txt
for grp in "gemm" "attention" "not gemm and not attention"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout -k 5 420 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" -p no:cacheprovider 2>&1 | tail -60 > "gpurun_out/kernels_${name}.log"
  echo "== $grp: exit ${PIPESTATUS[0]}"; tail -40 "gpurun_out/kernels_${name}.log"
done
