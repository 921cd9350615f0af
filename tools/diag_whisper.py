"""Where does the logit error of the Whisper-family path come from on a full 128-window forward call?

    python tools/diag_whisper.py

Runs the first 128-window batch of tests/golden/whisper_config2.npz through the engine, then
  * per-window max |logit error| against the reference's golden logits (does it grow along the recurrence?),
  * the LSTM + heads of the oracle (fp32, CPU) fed with the ENGINE's own encoder output ``mix``: the difference to
    the engine's logits is the error of the LSTM path alone; the difference to the golden is what the encoder
    contributes.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import segma_oracle as O  # noqa: E402
from segma_b200 import synth  # noqa: E402
from segma_b200.config import make_config  # noqa: E402
from segma_b200.encoders import MultiLabelEncoder  # noqa: E402
from segma_b200.models import Models  # noqa: E402


def main():
    g = np.load(Path(__file__).resolve().parent.parent / "tests" / "golden" / "whisper_config2.npz")
    n, audio_seed, bs = (int(v) for v in g["meta"])
    labels = synth.DEFAULT_LABELS
    sd = synth.surgical_hydra_state_dict(synth.WHISPER_SMALL, seed=0)
    model = Models["surgical_hydra"].from_state_dict(sd, MultiLabelEncoder(list(labels)), make_config("surgical_hydra"))
    eng = model.engine
    pcm = torch.from_numpy(synth.synth_audio(n, audio_seed)).cuda()
    nw = 128
    logits = torch.zeros((nw * 199, 4), device="cuda")
    eng.forward_pcm(pcm, 0, nw, 64000, 63680, logits, 0, 199, 199)
    torch.cuda.synchronize()
    got = logits.cpu()
    ref = torch.from_numpy(g["logits"])[: nw * 199]
    err = (got - ref).abs().view(nw, -1)
    print("logit std", ref.std().item(), "max|err|", err.max().item(), "mean|err|", err.mean().item())
    print("per-window max|err| (groups of 8 windows):", [round(float(v), 4) for v in err.view(16, -1).max(dim=1).values])
    print("per-window mean|err| (groups of 8 windows):", [round(float(v), 5) for v in err.view(16, -1).mean(dim=1)])
    w, f = divmod(int(err.view(-1, 4).max(dim=1).values.argmax()), 199)
    print("worst decision at window", w, "frame", f)
    mix = eng._ws["mix"][:nw].float().cpu()  # (nw, 199, d)
    torch.set_num_threads(16)
    with torch.inference_mode():
        from_mix = O._heads(sd, O.lstm_seq_first(sd, mix), labels).reshape(-1, 4)
    e_lstm = (got - from_mix).abs()
    e_enc = (from_mix - ref).abs()
    print(f"LSTM path alone (engine logits vs fp32 LSTM on the engine's mix): max {e_lstm.max().item():.4g} mean {e_lstm.mean().item():.4g}")
    print(f"encoder contribution (fp32 LSTM on the engine's mix vs golden): max {e_enc.max().item():.4g} mean {e_enc.mean().item():.4g}")
    # sensitivity of the fp32 LSTM to its input: perturb mix by fp16 rounding only
    with torch.inference_mode():
        from_mix16 = O._heads(sd, O.lstm_seq_first(sd, mix.half().float()), labels).reshape(-1, 4)
    e16 = (from_mix16 - from_mix).abs()
    print(f"fp32 LSTM, mix rounded to fp16 vs not: max {e16.max().item():.4g} mean {e16.mean().item():.4g}")
    print("mix rms", mix.pow(2).mean().sqrt().item())
    q = torch.quantile(err.reshape(-1)[:: 4].double(), torch.tensor([0.5, 0.99, 0.999, 0.9999], dtype=torch.float64))
    print("logit |err| quantiles 50 / 99 / 99.9 / 99.99 %:", [round(float(v), 5) for v in q])
    omix_p = Path(__file__).resolve().parent.parent / "tmp_diag" / "mix_oracle.pt"
    if omix_p.exists():  # the oracle's encoder output for the same batch (computed on the CPU beforehand)
        omix = torch.load(omix_p)
        dm = mix - omix
        rms = omix.pow(2).mean().sqrt().item()
        common = dm.mean(0, keepdim=True)
        print(f"mix error: rms/rms {dm.pow(2).mean().sqrt().item() / rms:.3e} max/rms {dm.abs().max().item() / rms:.3e}; "
              f"part common to all windows rms/rms {common.pow(2).mean().sqrt().item() / rms:.3e}, "
              f"window-varying part {(dm - common).pow(2).mean().sqrt().item() / rms:.3e}")

        def logits_of(m):
            with torch.inference_mode():
                return O._heads(sd, O.lstm_seq_first(sd, m), labels).reshape(-1, 4)

        base = logits_of(omix)
        for name, m in (("common error only", omix + common), ("window-varying error only", omix + (dm - common))):
            e = (logits_of(m) - base).abs()
            print(f"fp32 LSTM on oracle mix + {name}: max {e.max().item():.4g} mean {e.mean().item():.4g}")


if __name__ == "__main__":
    main()
