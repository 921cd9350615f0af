"""TEST INFRASTRUCTURE ONLY -- how fast is the oracle port next to the code it stands in for?

    python oracle/calibrate_port.py

bench.py's CPU arm times the oracle (``kind: "port"``) because the reference cannot travel to the GPU box.  This
script runs the port and the reference's own ``SurgicalHydra`` class (through oracle/ref_shim.py; Whisper-small dims,
the weights bench.py uses) on the same 8 windows, alternating, and writes ``oracle/calibration.json``:
``port_over_reference`` = best port time / best reference time.  bench.py prints it as ``cpu_baseline.calibration``
so that the headline ratio can be read against the reference itself; tests/test_oracle_vs_reference.py asserts the
port is not slower than the reference by more than 10 %.
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def measure(n_windows: int = 8, rounds: int = 3, threads: int = 8) -> dict:
    from oracle import ref_shim, segma_oracle as O
    from oracle.make_golden import LABELS, LSTMConfig, Models, MultiLabelEncoder, SurgicalHydraConfig, _whisper_dir
    from segma_b200 import synth

    torch.set_num_threads(threads)
    dims = synth.WHISPER_SMALL
    sd = synth.surgical_hydra_state_dict(dims, seed=0)
    cfg = ref_shim.make_config("surgical_hydra", SurgicalHydraConfig(encoder=_whisper_dir(dims), encoder_layers=[], reduction="weighted",
                                                                      lstm=LSTMConfig(128, 2, True, 0.5), classifier=256))
    ref_model = Models["surgical_hydra"](MultiLabelEncoder(list(LABELS)), cfg).eval()
    ref_model.load_state_dict(sd, strict=True)
    feats = torch.stack([O.whisper_logmel(torch.from_numpy(synth.synth_audio(64000, s))) for s in range(n_windows)])
    t_ref, t_port, diff = [], [], 0.0
    with torch.inference_mode():
        ref_model(feats[:1])
        O.surgical_hydra_forward(sd, feats[:1], LABELS)
        for _ in range(rounds):
            t0 = time.perf_counter()
            a = ref_model(feats)
            t_ref.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            b = O.surgical_hydra_forward(sd, feats, LABELS)
            t_port.append(time.perf_counter() - t0)
            diff = max(diff, (a - b).abs().max().item())
    return {"port_over_reference": min(t_port) / min(t_ref), "port_s": min(t_port), "reference_s": min(t_ref),
            "windows": n_windows, "threads": threads, "rounds": rounds, "max_abs_logit_diff": diff,
            "what": "oracle.surgical_hydra_forward vs the reference's SurgicalHydra.forward (Whisper-small dims), "
                    "best of alternating rounds on the build container's CPU"}


if __name__ == "__main__":
    res = measure()
    (ROOT / "oracle" / "calibration.json").write_text(json.dumps(res, indent=1) + "\n")
    print(json.dumps(res, indent=1))
