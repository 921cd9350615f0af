N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02l_bench_${N}gpu.json 2> gpurun_out/r02l_bench_${N}gpu.err
tail -2 gpurun_out/r02l_bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 2 --warmup 1 --workload corpus > gpurun_out/r02l_corpus_${N}gpu.json 2> gpurun_out/r02l_corpus_${N}gpu.err
tail -2 gpurun_out/r02l_corpus_${N}gpu.err
python - <<PY
import json
for f in ("r02l_bench_${N}gpu","r02l_corpus_${N}gpu"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["scaling"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
