"""Multi-GPU check of the sharded drivers on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/multi_gpu_check.py

Every rank builds the same small corpus of wav files (seeded), then
  1. ``run_inference_on_audios(..., shard=(rank, world))``: each rank writes the RTTMs of its own files
     (``distributed.assign_files``, longest first); together they cover the corpus exactly once;
  2. ``infer_corpus(..., shard=(rank, world))``: work is partitioned by file and -- for the one long file -- by window batch
     (``geometry.plan_work_units``); the table every rank gets from the final NCCL all-gather (+ merge of the pieces)
     equals the table a single process computes for the whole corpus, and equals the RTTMs of step 1.
"""
import os
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from segma_b200 import synth  # noqa: E402
from segma_b200.config import make_config  # noqa: E402
from segma_b200.distributed import assign_files, init_from_env  # noqa: E402
from segma_b200.inference import infer_corpus, run_inference_on_audios  # noqa: E402
from segma_b200.io import write_wav  # noqa: E402
from segma_b200.models import Models  # noqa: E402
from segma_b200.encoders import MultiLabelEncoder  # noqa: E402


def main():
    rank, world, local = init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    labels = synth.DEFAULT_LABELS
    kind = sys.argv[1] if len(sys.argv) > 1 else "surgical_hydra"
    root = Path(tempfile.gettempdir()) / f"segma_mg_{os.environ.get('MASTER_PORT', '0')}"
    wavs = root / "wav"
    lens = [64000 * 3 + 5000, 300, 64000, 70_000, 63680 * 6 + 9000, 5000, 63680 * 2 + 64000, 12_345, 64000 * 9, 40_000,
            63680 * 44 + 64000 + 9000]  # the last one is long enough to be cut into window-batch ranges across the ranks
    if rank == 0:
        wavs.mkdir(parents=True, exist_ok=True)
        for i, n in enumerate(lens):
            write_wav(wavs / f"f{i:02d}.wav", synth.synth_audio(n, 100 + i), subtype="int16")
        cfg = make_config(kind)
        cfg.save(root / "config.yml")
        sd = (synth.surgical_hydra_state_dict(synth.WHISPER_TEST, seed=3) if kind == "surgical_hydra"
              else synth.hubert_hydra_state_dict(synth.W2V2_TEST, seed=5))
        torch.save({"state_dict": sd}, root / "best.ckpt")
    if world > 1:
        dist.barrier()
    files = sorted(wavs.glob("*.wav"))
    out = root / f"out_w{world}"
    mine = run_inference_on_audios(config=root / "config.yml", uris=None, wavs=wavs, checkpoint=root / "best.ckpt", output=out,
                                   thresholds=None, batch_size=4, device=f"cuda:{local}", shard=(rank, world))
    sizes = [synth.synth_audio(1, 0).size * 0 + n for n in lens]
    assert [p.name for p in mine] == [files[i].name for i in assign_files(sizes, world)[rank]], "file assignment differs from assign_files"
    if world > 1:
        dist.barrier()
    rttms = sorted((out / "raw_rttm").glob("*.rttm"))
    assert [p.stem for p in rttms] == [p.stem for p in files], "the ranks together must cover every file exactly once"
    # final all-gather of the interval tables vs a single-process run of the whole corpus
    le = MultiLabelEncoder(list(labels))
    cfg = make_config(kind)
    blob = torch.load(root / "best.ckpt", map_location="cpu")
    model = Models[kind].from_state_dict(blob["state_dict"], le, cfg).to(dev)
    table = infer_corpus(files, model, cfg, batch_size=4, device=dev, shard=(rank, world)).cpu()
    whole = infer_corpus(files, model, cfg, batch_size=4, device=dev).cpu()
    assert torch.equal(table, whole), "gathered table differs from the single-process table"
    # ... and the RTTMs say the same
    n_lines = sum(len(p.read_text().splitlines()) for p in rttms)
    assert n_lines == whole.shape[0], (n_lines, whole.shape)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, int(table.shape[0]))
        assert len(set(gathered)) == 1
    if rank == 0:
        from segma_b200.geometry import assign_units, plan_work_units
        units = plan_work_units(sizes, world, 64000, 4, 63680, model.n_keep if model.family == "whisper" else 199)
        print(f"multi-GPU check ok: world {world}, {len(files)} files, {whole.shape[0]} intervals, model {kind}; "
              f"files per rank {[len(v) for v in assign_files(sizes, world)]}; work units per rank "
              f"{[len(v) for v in assign_units(units, world)]} ({sum(not u.whole_file for u in units)} pieces of cut files)")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
