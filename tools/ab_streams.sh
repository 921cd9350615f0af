#!/bin/bash
# A/B of SEGMA_STREAMS (window batches of one file on several streams / workspace slots), default bench workload
for v in 1 2 3 1 2 3; do
  SEGMA_STREAMS=$v python bench.py --steps 4 --warmup 3 --no-extra-workloads --no-cpu-baseline --no-side-kernels > gpurun_out/ab_st_$v.json 2> gpurun_out/ab_st_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_st_$v.json')); print('SEGMA_STREAMS=$v', round(d['value'],4), 'audio-h/s', round(d['ms_per_step'],1), 'ms/step e2e', round(d['e2e']['value'],4), d['clocks']['sm_mhz'])"
done
