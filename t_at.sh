for m in 0 0x2222 0x2A2A 0xAAAA; do
  echo "== SEGMA_ATTN_POLY=$m"
  SEGMA_ATTN_POLY=$m python -m pytest tests/test_kernels_gpu.py -q -k "attention" 2>&1 | tail -2
  SEGMA_ATTN_POLY=$m python tools/run_attention.py 128 1500 2>&1 | tail -2
done
