"""Per-stage comparison of the wav2vec2-family CUDA path with the oracle (diagnosis of logit error / label flips).

    python tools/diag_w2v2.py [hubert|wavlm] [n_windows]

Prints, for every stage (conv0..6, projection, positional conv, each encoder layer), the error of the CUDA
path's output against the fp32 oracle relative to the stage's rms, then the logit error and the label agreement.
Runs on the GPU box; the oracle is test infrastructure and is only the checker here.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from oracle import segma_oracle as O  # noqa: E402
from segma_b200 import synth  # noqa: E402
from segma_b200.config import make_config  # noqa: E402
from segma_b200.encoders import MultiLabelEncoder  # noqa: E402
from segma_b200.models import Models  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "hubert"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    dims, seed = (synth.WAVLM_BASE, 6) if which == "wavlm" else (synth.HUBERT_BASE, 5)
    labels = synth.DEFAULT_LABELS
    sd = synth.hubert_hydra_state_dict(dims, seed=seed)
    model = Models["surgical_hubert_hydra"].from_state_dict(sd, MultiLabelEncoder(list(labels)), make_config("surgical_hubert_hydra"))
    wav = torch.stack([torch.from_numpy(synth.synth_audio(64000, 100 + s)) for s in range(n)])
    eng = model.engine
    eng.trace = []
    got = model(wav).cpu().reshape(-1, 4)
    torch.cuda.synchronize()
    dev_trace = [(k, v.cpu()) for k, v in eng.trace]
    eng.trace = None
    torch.set_num_threads(16)
    ref_trace = []
    with torch.inference_mode():
        ref = O.hubert_hydra_forward(sd, wav, labels, trace=ref_trace).reshape(-1, 4)
    rt = dict(ref_trace)
    print(f"{which}: {n} windows")
    for name, g in dev_trace:
        r = rt[name]
        err = (g - r).abs()
        rms = r.pow(2).mean().sqrt().item()
        print(f"  {name:12s} rms {rms:9.4f}  max|err|/rms {err.max().item() / rms:9.2e}  mean|err|/rms {err.mean().item() / rms:9.2e}")
    err = (got - ref).abs()
    flips = ((got > 0) != (ref > 0))
    print(f"  logits: std {ref.std().item():.3f} max|err| {err.max().item():.3e} mean|err| {err.mean().item():.3e} "
          f"agreement {1 - flips.float().mean().item():.5f} ({int(flips.sum())} flips of {flips.numel()})")
    if flips.any():
        print("  |ref logit| at the flipped decisions:", [round(v, 5) for v in ref[flips].abs().tolist()][:40])
        print("  fraction of |ref logit| below max|err|:", (ref.abs() < err.max()).float().mean().item())


if __name__ == "__main__":
    main()
