#!/bin/bash
# A/B of cross-file window packing for a wav2vec2-family model on a corpus of short clips (median 15 s)
for v in 1 0; do
  SEGMA_PACK_FILES=$v python bench.py --workload corpus --corpus-model hubert --corpus-median-s 15 --corpus-files 1024 --steps 2 --warmup 1 > gpurun_out/ab_pack_$v.json 2> gpurun_out/ab_pack_$v.err
  tail -1 gpurun_out/ab_pack_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_pack_$v.json')); print('SEGMA_PACK_FILES=$v', d['value'], 'audio-h/s', d['ms_per_step'], 'ms/step e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['config']['workload'][:90])"
done
