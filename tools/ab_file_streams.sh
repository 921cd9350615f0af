#!/bin/bash
# A/B of SEGMA_FILE_STREAMS (short files of a Whisper-family corpus taking turns on several streams)
python -m pytest tests/test_e2e_gpu.py -q -k "packed" 2>&1 | tail -3
for med in 15 150; do
for v in 4 1; do
  SEGMA_FILE_STREAMS=$v python bench.py --workload corpus --corpus-median-s $med --corpus-files 256 --steps 2 --warmup 1 > gpurun_out/ab_fs_${med}_$v.json 2> gpurun_out/ab_fs_${med}_$v.err
  tail -1 gpurun_out/ab_fs_${med}_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_fs_${med}_$v.json')); print('median $med s SEGMA_FILE_STREAMS=$v', round(d['value'],3), 'audio-h/s', round(d['ms_per_step'],1), 'ms/step e2e', round(d['e2e']['value'],3), d['config']['workload'][:60])"
done; done
