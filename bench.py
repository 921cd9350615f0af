#!/usr/bin/env python
"""Headline benchmark: audio-hours/sec of the sliding-window inference path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (`config.workload`): Whisper-small-dims encoder + segma LSTM/linear heads (`surgical_hydra`),
80-bin log-mel, one 1 h synthetic 16 kHz file per step and per GPU, 4 s windows at the reference's step
(905 forward windows, 179 999 frames), batch 128, default 0.5 thresholds, intervals decoded on the device.
A step = the whole hot path over one file: window -> log-mel -> encoder -> LSTM -> heads -> logits on the
file timeline -> threshold + run-length decode -> interval table.  Nothing is read back between steps; after the
last step the row counts are read once and (N > 1) the tables are all-gathered once over NCCL, inside the timed region.

  value  : device-timed (CUDA events), PCM already resident in HBM.
  e2e    : the same through the public per-file API (what `infer_file` / `run_inference_on_audios` run) from pinned
           host PCM: every step's H2D copy of its file and D2H read of its interval table inside the timed region.
  roofline: the dominant kernel (tcgen05 GEMM / implicit conv), algorithmic FLOPs / CUDA-event time of its
           launches inside the timed region, against the measured sustained bf16 cuBLAS peak.
  cpu_baseline: the oracle (reference arithmetic in torch fp32) on this box's host cores on a bounded
           sample of the same workload.
  gpu_eager_baseline: the reference's own modules (transformers WhisperEncoder, nn.LSTM, nn.Linear) in torch eager on
           the same GPU -- the library kernels (cuBLASLt / cuDNN / SDPA / cuFFT) the hand-written ones replace.
  workloads: HuBERT-base / WavLM-base+ dims (BASELINE configs 1 and 3 models), 3 steps of 1 h each.
`--workload corpus`: BASELINE config 4 in miniature (skewed file durations, file-sharded, strong scaling).
`--impl reference` times the CPU arm alone (all host threads), same metric / unit / config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

HOUR_SAMPLES = 57_600_000
WIN, STEP_SAMPLES, BATCH = 64_000, 63_680, 128
LABELS = ("KCHI", "OCH", "MAL", "FEM")
METRIC, UNIT = "audio_hours_per_sec", "audio-h/s"


WORKLOADS = {
    # name: (model kind, description, GFLOP per 4 s window)
    "whisper": ("surgical_hydra", "whisper-small-dims surgical_hydra (12x768, LSTM 2x128 bidir, 4 heads), 80-bin log-mel", 344.7),
    "hubert": ("surgical_hubert_hydra", "HuBERT-base-dims surgical_hubert_hydra (7-layer conv front end, 12x768, 4 heads)", 56.9),
    "wavlm": ("surgical_hubert_hydra", "WavLM-base+-dims encoder in surgical_hubert_hydra (gated relative-position bias)", 56.9),
}


def _config(n_gpus: int, audio_s: float, workload: str = "whisper") -> dict:
    kind, desc, _ = WORKLOADS[workload]
    return {
        "workload": f"{desc}, "
                    f"{audio_s / 3600:.4g} h synthetic 16 kHz audio per GPU per step, 4 s windows step 63680, batch 128",
        "weights": "random init (segma_b200.synth, seed 0), reference state_dict layout",
        "audio_seconds_per_gpu_step": audio_s,
        "window_batch": BATCH,
        "parallelism": f"files sharded over {n_gpus} GPU(s), interval all-gather" if n_gpus > 1 else "single GPU",
        "l2": "inputs larger than L2 (230 MB PCM per file, >3 GB activations per batch)",
    }


# ---- stdout carries exactly one JSON line -------------------------------------------------------------------
_REAL_STDOUT = None


def protect_stdout() -> None:
    """Libraries write banners to file descriptor 1 (NCCL prints its version there at NCCL_DEBUG=VERSION unless
    NCCL_DEBUG_FILE says otherwise; torchrun children inherit the descriptor): keep a private copy of stdout for the
    result line and point descriptor 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ---- clocks ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples: list[list[str]] = []
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                row = [c.strip() for c in out.strip().split(",")]
                if len(row) == 6:
                    self.samples.append(row)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(r[0]) for r in self.samples if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.samples if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---- CPU arm (oracle) --------------------------------------------------------------------------------------
def cpu_reference_run(n_windows: int, steps: int, warmup: int, seed: int = 0):
    """Times the oracle (reference arithmetic, torch fp32, all host threads) on `n_windows` 4 s windows per
    step: log-mel hook per window -> SurgicalHydra forward (one batch) -> thresholds -> intervals."""
    import torch

    from oracle import segma_oracle as O
    from segma_b200 import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.surgical_hydra_state_dict(synth.WHISPER_SMALL, seed=0)
    n = STEP_SAMPLES * (n_windows - 1) + WIN
    pcm = torch.from_numpy(synth.synth_audio(n, seed))

    def one_step():
        logits = O.apply_model_on_audio(pcm, lambda f: O.surgical_hydra_forward(sd, f, LABELS), len(LABELS),
                                        batch_size=BATCH, whisper=True)
        mask = O.apply_thresholds(logits, [0.5] * len(LABELS))
        return O.create_intervals(mask.numpy(), LABELS)

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    audio_s = n / 16_000
    return {"audio_s": audio_s, "sec_per_step": dt, "value": audio_s / 3600.0 / dt, "cores": cores,
            "sample": f"{n_windows} windows ({audio_s:.1f} s of audio) per step, {steps} step(s), torch fp32 on "
                      f"{cores} host threads; linear in audio length"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_run(args.ref_windows, max(1, args.steps), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["sec_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args.gpus, res["audio_s"]),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"],
                         "calibration": _calibration()},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- torch-eager-on-B200 arm: the reference's own module stack on the GPU ---------------
def eager_gpu_baseline(dev, sd, n_windows: int = 128) -> dict:
    """The reference path as torch 2.11 eager dispatches it on this GPU: transformers' ``WhisperEncoder`` (SDPA),
    ``nn.LSTM`` read sequence-first (the reference never sets batch_first), per-label ``nn.Linear`` heads -- composed
    as /root/reference/src/segma/models/whisper/surgical_hydra.py:80-109 composes them -- with the log-mel hook
    restated in torch on the device (torch.stft -> mel -> log10 -> clamp, feature_extraction_whisper.py:135-164).
    cuFFT / cuDNN / cuBLASLt / flash SDPA: the library kernels the hand-written ones have to beat.  One 128-window
    batch per step (the reference's default), 1 warm-up + 2 timed steps per precision."""
    import numpy as np
    import torch
    from torch import nn
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder

    from segma_b200 import ops, synth

    cfg = WhisperConfig(d_model=768, encoder_layers=12, encoder_attention_heads=12, encoder_ffn_dim=3072, decoder_layers=1,
                        decoder_attention_heads=2, decoder_ffn_dim=64, vocab_size=100)
    enc = WhisperEncoder(cfg).eval()
    enc.load_state_dict({k[len("w_encoder."):]: v for k, v in sd.items() if k.startswith("w_encoder.")}, strict=True)
    lstm = nn.LSTM(input_size=768, hidden_size=128, num_layers=2, bidirectional=True).eval()
    lstm.load_state_dict({k[len("lstm_shared."):]: v for k, v in sd.items() if k.startswith("lstm_shared.")}, strict=True)
    heads = nn.ModuleList([nn.Linear(256, 1) for _ in LABELS]).eval()
    for h, lab in zip(heads, LABELS):
        h.load_state_dict({"weight": sd[f"task_heads.linear_head_{lab}.weight"], "bias": sd[f"task_heads.linear_head_{lab}.bias"]})
    enc, lstm, heads = enc.to(dev), lstm.to(dev), heads.to(dev)
    lw = torch.softmax(sd["layer_weights"], 0).to(dev)
    mel = torch.from_numpy(np.ascontiguousarray(ops.mel_filters().T)).to(dev)  # (80, 201), same filterbank values
    hann = torch.hann_window(400, device=dev)
    n = STEP_SAMPLES * (n_windows - 1) + WIN
    pcm = torch.from_numpy(synth.synth_audio(n, 0)).to(dev)

    def step():
        wins = pcm.unfold(0, WIN, STEP_SAMPLES)  # (n_windows, 64000), inference.py:148-152
        x = torch.zeros((n_windows, 480_000), device=dev)
        x[:, :WIN] = wins
        st = torch.stft(x, 400, 160, window=hann, return_complex=True)
        spec = mel @ (st[..., :-1].abs() ** 2)
        log_spec = torch.clamp(spec, min=1e-10).log10()
        log_spec = torch.maximum(log_spec, log_spec.amax(dim=(1, 2), keepdim=True) - 8.0)
        feats = (log_spec + 4.0) / 4.0
        hs = enc(feats, output_hidden_states=True).hidden_states[1:]
        mix = torch.einsum("l,l...->...", lw.to(hs[0].dtype), torch.stack(list(hs), dim=0))
        out, _ = lstm(mix)  # sequence axis = windows
        out = out[:, :199]
        logits = torch.stack([h(out) for h in heads], dim=-1).reshape(-1, len(LABELS))
        return (logits.float().sigmoid() > 0.5)

    res = {"windows_per_step": n_windows, "audio_s_per_step": n / 16_000,
           "what": "transformers WhisperEncoder (sdpa) + nn.LSTM + nn.Linear heads + torch.stft log-mel, torch eager, "
                   "weights and audio as in the main arm; thresholded on the device, no interval decode"}
    for name, ctx in (("fp32", None), ("bf16_autocast", torch.bfloat16), ("fp16_autocast", torch.float16)):
        try:
            with torch.inference_mode():
                def run():
                    if ctx is None:
                        return step()
                    with torch.autocast("cuda", dtype=ctx):
                        return step()
                run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(2):
                    run()
                e1.record()
                torch.cuda.synchronize()
            sec = e0.elapsed_time(e1) / 2e3
            res[name] = {"value": n / 16_000 / 3600.0 / sec, "unit": UNIT, "ms_per_128_windows": sec * 1e3}
        except Exception as e:  # noqa: BLE001  (an arm that cannot run is reported, not hidden)
            res[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
    del enc, lstm, heads
    torch.cuda.empty_cache()
    return res


# ---- GPU arm ---------------------------------------------------------------------------------------------
def _peaks():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:  # noqa: BLE001
        return {}


def _traffic():
    """DRAM bytes per launch of the dominant kernel from this round's ``ncu --set full`` capture (ncu cannot run inside
    the bench; tools/ncu_full_summary.py writes the file from the capture of the same command)."""
    try:
        return json.loads((ROOT / "profiles" / "r02z_gemm_traffic.json").read_text())
    except Exception:  # noqa: BLE001
        return None


def measure_model_workload(args, workload: str, steps: int, warmup: int, rank: int, world: int, dev, with_e2e: bool = True):
    """One model workload: K timed steps (one file of ``args.hours`` h per GPU per step) + the final exchange."""
    import torch
    import torch.distributed as dist

    from segma_b200 import ops, synth
    from segma_b200.config import make_config
    from segma_b200.distributed import gather_corpus_tables
    from segma_b200.encoders import MultiLabelEncoder
    from segma_b200.geometry import INFERENCE_SETTINGS
    from segma_b200.inference import MAX_FILES_IN_FLIGHT, _FileJob, apply_model_on_audio, default_thresholds
    from segma_b200.models import Models
    from segma_b200.thresholds import logit_cut

    n_samples = int(args.hours * HOUR_SAMPLES)
    audio_s = n_samples / 16_000
    le = MultiLabelEncoder(list(LABELS))
    kind, _, gflop_per_window = WORKLOADS[workload]
    cfg = make_config(kind)
    if workload == "whisper":
        sd = synth.surgical_hydra_state_dict(synth.WHISPER_SMALL, seed=0)
    else:
        sd = synth.hubert_hydra_state_dict(synth.WAVLM_BASE if workload == "wavlm" else synth.HUBERT_BASE, seed=0)
    model = Models[kind].from_state_dict(sd, le, cfg).to(dev)
    cuts = [logit_cut(0.5)] * len(LABELS)
    thr = default_thresholds(le)
    host_pcm = torch.from_numpy(synth.synth_audio(n_samples, seed=rank)).pin_memory()
    dev_pcm = host_pcm.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(src, k, to_host):
        """k files through the path with nothing read back in between, then the single end-of-run exchange: counts to
        the host once, tables compacted and all-gathered once (NCCL for world > 1)."""
        tables, counts = [], []
        for _ in range(k):
            logits = apply_model_on_audio(src, model, INFERENCE_SETTINGS, dev, batch_size=BATCH)
            t, c = ops.decode_intervals_async(logits, cuts, mode=ops.DECODE_LOGIT)
            tables.append(t)
            counts.append(c)
        full = gather_corpus_tables([rank * k + i for i in range(k)], tables, counts, device=dev, gather=world > 1)
        return full.cpu() if to_host else full

    def run_e2e(src, k, to_host):
        """The public per-file API from pinned host PCM: every step stages its file over PCIe behind the compute of the
        previous batches and brings its interval table (and row count) back to the host on a side stream, where it is
        turned into the reference's ``(start, end, label)`` list; up to MAX_FILES_IN_FLIGHT files are queued before
        the host waits for the oldest.  For world > 1 the final all-gather of the device tables follows."""
        jobs, d2h = [], 0
        for i in range(k):
            jobs.append(_FileJob(src, model, cfg, BATCH, dev, thr, False, None))
            if i >= MAX_FILES_IN_FLIGHT:
                jobs[i - MAX_FILES_IN_FLIGHT].finish(None)
        for j in jobs[max(0, k - MAX_FILES_IN_FLIGHT):]:
            j.finish(None)
        d2h = sum(j.table.numel() * 4 + 4 for j in jobs)  # the worst-case sized table + its row count, per file
        if world > 1:
            gather_corpus_tables([rank * k + i for i in range(k)], [j.table for j in jobs], [j.count for j in jobs],
                                 device=dev, gather=True)
        return d2h

    def timed(src, k, to_host):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = (run_e2e if to_host else run_steps)(src, k, to_host)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item() / 1e3, ms[1].item() / 1e3, out

    run_steps(dev_pcm, warmup, False)
    if with_e2e:
        run_e2e(host_pcm, 1, True)
    with ClockSampler(dev.index or 0) as clocks:
        ops.stats.reset()
        dev_s, _, table = timed(dev_pcm, steps, False)
        launches = ops.stats.launches
        e2e_wall, d2h_bytes = None, 0
        if with_e2e:
            _, e2e_wall, d2h_bytes = timed(host_pcm, steps, True)

    # roofline of the dominant kernel: a separately timed pass with an event pair around every launch
    ops.stats.reset()
    ops.stats.profile = True
    barrier()
    run_steps(dev_pcm, 1, False)
    torch.cuda.synchronize()
    ops.stats.profile = False
    breakdown = ops.stats.breakdown()
    gemms = [v for k, v in breakdown.items() if k.startswith("segma_gemm_f16")]
    g_ms, g_flops, n_gemm = sum(v["ms"] for v in gemms), sum(v["work"] for v in gemms), sum(v["calls"] for v in gemms)
    executed = sum(v["work"] for k, v in breakdown.items() if k.startswith(("segma_gemm_f16", "segma_attention")))
    peaks = _peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    achieved_tf = g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    n_windows = 905 * (n_samples / HOUR_SAMPLES)
    traffic = _traffic() if workload == "whisper" else None
    res = {
        "value": world * (audio_s / 3600.0) * steps / dev_s,
        "ms_per_step": dev_s / steps * 1e3,
        "audio_s": audio_s,
        "launches": launches,
        "intervals_per_step": int(table.shape[0]) // max(steps * world, 1),
        # tensor-core FLOPs the launches of one step actually execute (last-layer pruning, tile padding of grouped
        # convolutions excluded) and the reference's algorithmic count for the same windows, both per GPU
        "model_tflops_per_gpu": executed / (dev_s / steps) / 1e12,
        "executed_gflop_per_window": executed / n_windows / 1e9,
        "algorithmic_gflop_per_window": gflop_per_window,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc5_kernel (tcgen05 GEMM + implicit conv)",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "traffic": traffic["bytes_per_launch"] if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None,
                     "launches": n_gemm, "avg_launch_ms": g_ms / max(n_gemm, 1),
                     "share_of_step": (g_ms / 1e3) / (dev_s / steps), "peak_source": peak_src},
        "clocks": clocks.summary(),
        "breakdown_ms_per_step": {k: [round(v["ms"], 3), round(v["work"] / max(v["ms"], 1e-9) / 1e9, 1)]
                                  for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])},
    }
    if with_e2e:
        res["e2e"] = {"value": world * (audio_s / 3600.0) * steps / e2e_wall, "unit": UNIT,
                      "h2d_bytes_per_step": host_pcm.numel() * 4,
                      "d2h_bytes_per_step": d2h_bytes // steps}
    return res, dict(model=model, sd=sd, dev_pcm=dev_pcm, cuts=cuts)


def measure_side_kernels(dev, ctx) -> dict:
    """HBM-bound kernels timed alone against the measured copy bandwidth (burst peak): the two front ends on a batch
    larger than L2 and threshold + run-length decode on >= 200 h of logits."""
    import torch

    from segma_b200 import ops, synth

    hbm_peak = float(_peaks().get("hbm_gbs", 6650.0))
    dev_pcm, cuts = ctx["dev_pcm"], ctx["cuts"]

    def sm_clock():
        try:
            out = subprocess.run(["nvidia-smi", f"--id={dev.index or 0}", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout
            return float(out.strip().splitlines()[0])
        except Exception:  # noqa: BLE001
            return None

    clocks_seen = {}

    def time_kernel(fn, iters=20, what=None):
        """A kernel timed ALONE: the power governor moves the SM clock in steps tens of milliseconds apart, so a
        measurement that follows the 1 kW model step directly would run at that step's clock (1.39-1.45 GHz) instead of
        the kernel's own (`tools/logmel_clock_check.py`: the log-mel kernel takes 0.61-0.73 us per window right after
        3 s of GEMMs and 0.49 from half a second later on, also over 200 launches in a row).  Hence a 1 s pause, a
        warm-up, and at least 0.25 s of back-to-back launches; the SM clock at the end is reported."""
        torch.cuda.synchronize()
        time.sleep(1.0)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        total, n = 0.0, 0
        while total < 0.25 and n < 50 * iters:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1) * 1e-3
            n += iters
        if what:
            clocks_seen[what] = sm_clock()
        return total / n

    side = {}
    lib = ops._lib()
    st = torch.cuda.current_stream().cuda_stream
    n_w = 1024  # 1024 windows: 262 MB in + 983 MB out, larger than L2
    span = STEP_SAMPLES * (n_w - 1) + WIN
    pcm_side = dev_pcm[:span] if dev_pcm.numel() >= span else torch.randn(span, device=dev) * 0.1
    out_f32 = torch.empty((n_w, 80, 3000), dtype=torch.float32, device=dev)
    scratch = torch.empty(ops.logmel_scratch_bytes(n_w, WIN), dtype=torch.uint8, device=dev)
    t_mel = time_kernel(lambda: lib.segma_logmel(pcm_side.data_ptr(), pcm_side.numel(), n_w, WIN, STEP_SAMPLES,
                                                 out_f32.data_ptr(), None, scratch.data_ptr(), st), what="logmel")
    mel_bytes = n_w * (4 * WIN + 4 * 80 * 3000)
    side["logmel"] = {"bound": "hbm", "achieved": mel_bytes / t_mel / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": mel_bytes / t_mel / 1e9 / hbm_peak, "us_per_window": t_mel / n_w * 1e6,
                      "algorithmic_bytes_per_window": 4 * WIN + 4 * 80 * 3000, "sm_mhz": clocks_seen.get("logmel")}
    del out_f32, scratch
    # wav2vec2 / HuBERT / WavLM front end, layer 0: conv(k=10, s=5) + GroupNorm + GELU, 4*64000 B read and the
    # 2*12799*512 B fp16 activation written per window (SURVEY.md 8d counts the front end's HBM-bound part)
    n_l0 = 256
    span0 = STEP_SAMPLES * (n_l0 - 1) + WIN
    sd0 = synth.hubert_hydra_state_dict(synth.HUBERT_BASE, seed=0)
    fe = "wav2vec2.feature_extractor.conv_layers.0."
    w0 = sd0[fe + "conv.weight"].reshape(512, 10).contiguous().to(dev)
    g0, b0 = sd0[fe + "layer_norm.weight"].to(dev), sd0[fe + "layer_norm.bias"].to(dev)
    rows0 = (WIN - 10) // 5 + 1
    act = torch.zeros((n_l0, rows0 + (rows0 & 1), 512), dtype=torch.float16, device=dev)
    ss = torch.empty((n_l0, 512, 2), dtype=torch.float32, device=dev)
    t_l0 = time_kernel(lambda: ops.w2v2_layer0(dev_pcm[:span0], n_l0, WIN, STEP_SAMPLES, w0, g0, b0, ss, act), iters=10, what="l0")
    l0_bytes = n_l0 * (4 * WIN + 2 * rows0 * 512)
    side["w2v2_layer0"] = {"bound": "hbm", "achieved": l0_bytes / t_l0 / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": l0_bytes / t_l0 / 1e9 / hbm_peak, "us_per_window": t_l0 / n_l0 * 1e6,
                           "algorithmic_bytes_per_window": 4 * WIN + 2 * rows0 * 512, "sm_mhz": clocks_seen.get("l0")}
    del act, ss

    def speech_like(n_fr):
        # runs of ~1 s per label (the reference's synthetic annotations are 0.2-3 s long); built in chunks of 100 h
        parts = []
        for lo in range(0, n_fr, 18_000_000):
            m = min(18_000_000, n_fr - lo)
            parts.append(torch.randn((m // 50, len(LABELS)), device=dev).repeat_interleave(50, dim=0)
                         + 0.05 * torch.randn((m, len(LABELS)), device=dev))
        return torch.cat(parts).contiguous()

    for name, hours, make in (
        # 1000 h of logits batched (2.9 GB), one file per hour
        ("decode", 1000, speech_like),
        # worst case: iid logits, one interval every ~4 frames per label (the 16 B/interval table dominates)
        ("decode_worst_case", 200, lambda n_fr: torch.randn((n_fr, len(LABELS)), device=dev)),
    ):
        n_fr = 180_000 * hours
        offs = [i * 180_000 for i in range(hours + 1)]
        big = make(n_fr)
        tbl = ops.decode_intervals(big, cuts, file_offsets=offs, mode=ops.DECODE_LOGIT)
        n_iv = int(tbl.shape[0])
        del tbl
        # device time of the C-ABI call (its kernels + two small offset uploads), CUDA events around it
        ops.stats.reset()
        ops.stats.profile = True
        for _ in range(5):
            ops.decode_intervals(big, cuts, file_offsets=offs, mode=ops.DECODE_LOGIT, capacity=n_iv)
        torch.cuda.synchronize()
        ops.stats.profile = False
        evs = [e0.elapsed_time(e1) for nm, e0, e1, _ in ops.stats.events if nm == "segma_decode_intervals"]
        t_dec = min(evs) * 1e-3
        dec_bytes = 4 * len(LABELS) * n_fr + 16 * n_iv
        side[name] = {"bound": "hbm", "achieved": dec_bytes / t_dec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": dec_bytes / t_dec / 1e9 / hbm_peak, "hours_of_logits": hours, "intervals": n_iv,
                      "ms": t_dec * 1e3,
                      "note": "device time of one segma_decode_intervals call (count, tile sums, scan, write kernels)"}
        del big
    return side


def measure_corpus(args, rank: int, world: int, dev):
    """BASELINE config 4 in miniature: a fixed corpus of files with skewed durations (10 s ... 1 h), sharded by file
    over the ranks (longest first), no exchange until the single final all-gather.  Strong scaling: the corpus does not
    grow with the number of GPUs."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from segma_b200 import ops, synth
    from segma_b200.config import make_config
    from segma_b200.encoders import MultiLabelEncoder
    from segma_b200.inference import infer_corpus
    from segma_b200.models import Models

    le = MultiLabelEncoder(list(LABELS))
    kind, desc, _ = WORKLOADS[args.corpus_model]
    cfg = make_config(kind)
    if args.corpus_model == "whisper":
        sd = synth.surgical_hydra_state_dict(synth.WHISPER_SMALL, seed=0)
    else:
        sd = synth.hubert_hydra_state_dict(synth.WAVLM_BASE if args.corpus_model == "wavlm" else synth.HUBERT_BASE, seed=0)
    model = Models[kind].from_state_dict(sd, le, cfg).to(dev)
    rng = np.random.default_rng(4)
    n_files = args.corpus_files
    if args.corpus_hours > 0:  # full-size config 4: draw files until the corpus holds that many hours
        n_files = max(64, int(args.corpus_hours * 3600.0 / (args.corpus_median_s * 1.7)))
    dur_s = np.clip(np.exp(rng.normal(np.log(args.corpus_median_s), 1.3, size=n_files)), min(10.0, args.corpus_median_s), 3600.0)
    if args.corpus_hours > 0:
        n_files = int(min(n_files, np.searchsorted(np.cumsum(dur_s), args.corpus_hours * 3600.0) + 1))
        dur_s = dur_s[:n_files]
    lens = (dur_s * 16_000).astype(np.int64)
    pool = torch.from_numpy(synth.synth_audio(HOUR_SAMPLES, seed=0)).pin_memory()
    offs = rng.integers(0, HOUR_SAMPLES - lens + 1)
    host_files = [pool[int(o): int(o) + int(n)] for o, n in zip(offs, lens)]
    dev_pool = pool.to(dev)
    dev_files = [dev_pool[int(o): int(o) + int(n)] for o, n in zip(offs, lens)]
    total_h = float(lens.sum()) / 16_000 / 3600.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(files, k, to_host):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            out = infer_corpus(files, model, cfg, BATCH, dev, shard=(rank, world), sizes=lens[:len(files)].tolist())
            if to_host:
                out = out.cpu()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item() / 1e3, ms[1].item() / 1e3, out

    # one warm-up pass; a corpus of hundreds of hours warms up on its first 256 files
    timed(dev_files[:256] if args.corpus_hours > 0 else dev_files, max(1, min(args.warmup, 1)), False)
    with ClockSampler(dev.index or 0) as clocks:
        ops.stats.reset()
        dev_s, _, table = timed(dev_files, args.steps, False)
        launches = ops.stats.launches
        _, e2e_wall, table_host = timed(host_files, args.steps, True)
    return {
        "value": total_h * args.steps / dev_s, "ms_per_step": dev_s / args.steps * 1e3, "launches": launches,
        "e2e": {"value": total_h * args.steps / e2e_wall, "unit": UNIT, "h2d_bytes_per_step": int(lens.sum()) * 4,
                "d2h_bytes_per_step": int(table_host.numel()) * 4},
        "clocks": clocks.summary(), "corpus_hours": total_h, "n_files": n_files, "intervals": int(table.shape[0]),
        "longest_file_s": float(dur_s.max()), "median_file_s": float(np.median(dur_s)), "model": desc,
    }


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from segma_b200 import ops
    from segma_b200.distributed import init_from_env

    # NCCL writes its version banner (NCCL_DEBUG=VERSION and up) to stdout unless told otherwise; stdout carries the
    # single JSON line of this script
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank, world, local = init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ops.device_check()

    if args.workload == "corpus":
        r = measure_corpus(args, rank, world, dev)
        if rank == 0:
            line = {
                "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16", "data": "synthetic",
                "config": {"workload": f"corpus: {r['n_files']} files, {r['corpus_hours']:.2f} h in total, durations log-normal "
                                       f"(median {r['median_file_s']:.0f} s, longest {r['longest_file_s']:.0f} s), "
                                       f"{r['model']}, 4 s windows step 63680, batch 128",
                           "parallelism": f"files sharded over {world} GPU(s), longest first; one final interval all-gather",
                           "l2": "a 230 MB PCM pool and > 3 GB of activations per 128-window batch: larger than L2"},
                "e2e": r["e2e"], "gpu_launches": r["launches"], "clocks": r["clocks"], "intervals_per_step": r["intervals"],
            }
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return

    res, ctx = measure_model_workload(args, args.workload, args.steps, args.warmup, rank, world, dev)
    extras, side, cpu, eager = {}, {}, None, None
    def attempt(what, fn):
        """The headline numbers are already measured: a side measurement that cannot run is reported, not fatal."""
        try:
            return fn()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] {what} failed: {type(e).__name__}: {e}", file=sys.stderr, flush=True)
            return {"error": f"{type(e).__name__}: {e}"[:300]}

    if world == 1:
        if not args.no_side_kernels:
            side = attempt("side kernels", lambda: measure_side_kernels(dev, ctx))
        if args.workload == "whisper" and not args.no_extra_workloads:
            if not args.no_eager_baseline:
                eager = attempt("torch-eager arm", lambda: eager_gpu_baseline(dev, ctx["sd"]))
            ctx.clear()
            torch.cuda.empty_cache()

            def extra(wl):
                r, c = measure_model_workload(args, wl, 3, 3, rank, world, dev, with_e2e=False)
                c.clear()
                torch.cuda.empty_cache()
                return {"workload": WORKLOADS[wl][1], "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
                        "steps": 3, "warmup": 3, "roofline": r["roofline"], "gpu_launches": r["launches"],
                        "model_tflops_per_gpu": r["model_tflops_per_gpu"],
                        "executed_gflop_per_window": r["executed_gflop_per_window"],
                        "breakdown_ms_per_step": r["breakdown_ms_per_step"]}

            for wl in ("hubert", "wavlm"):  # BASELINE configs 1 and 3 models, driver-visible: 3 steps of 1 h each
                extras[wl] = attempt(f"workload {wl}", lambda wl=wl: extra(wl))
        if not args.no_cpu_baseline and args.workload == "whisper":
            def cpu_arm():
                c = cpu_reference_run(args.ref_windows, 1, 0)
                return {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": "port", "sample": c["sample"],
                        "calibration": _calibration()}

            cpu = attempt("cpu baseline", cpu_arm)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16",
        "dtype_detail": "f16 tensor-core operands; f32 accumulation, residual stream, LayerNorm/softmax statistics, LSTM "
                        "(input projection and recurrence at fp32 level), logits",
        "data": "synthetic", "config": _config(world, res["audio_s"], args.workload),
        "realtime_factor_per_gpu": res["value"] * 3600.0 / world,
        "model_tflops_per_gpu": res["model_tflops_per_gpu"],
        "executed_gflop_per_window": res["executed_gflop_per_window"],
        "algorithmic_gflop_per_window": res["algorithmic_gflop_per_window"],
        "e2e": res["e2e"], "gpu_launches": res["launches"], "intervals_per_step": res["intervals_per_step"],
        "roofline": res["roofline"], "cpu_baseline": cpu, "gpu_eager_baseline": eager, "side_kernels": side,
        "workloads": extras, "clocks": res["clocks"], "breakdown_ms_per_step": res["breakdown_ms_per_step"],
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def _calibration():
    """Port / reference time ratio of the CPU arm, measured where the reference is mounted (oracle/calibrate_port.py)."""
    try:
        return json.loads((ROOT / "oracle" / "calibration.json").read_text())
    except Exception:  # noqa: BLE001
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="segma_b200", choices=["segma_b200", "reference"])
    ap.add_argument("--hours", type=float, default=1.0, help="audio hours per GPU per step")
    ap.add_argument("--ref-windows", type=int, default=16, help="windows per step of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="whisper", choices=sorted(WORKLOADS) + ["corpus"],
                    help="whisper = BASELINE config 2 (the headline, also reports hubert / wavlm = configs 1 and 3 under "
                         "'workloads'); corpus = config 4 in miniature (file-sharded, strong scaling)")
    ap.add_argument("--corpus-files", type=int, default=256, help="files in the corpus workload")
    ap.add_argument("--corpus-hours", type=float, default=0.0,
                    help="size the corpus by its total duration instead (1000 = BASELINE config 4 as stated)")
    ap.add_argument("--corpus-model", default="whisper", choices=sorted(WORKLOADS), help="model of the corpus workload")
    ap.add_argument("--corpus-median-s", type=float, default=150.0, help="median file duration of the corpus workload")
    ap.add_argument("--no-side-kernels", action="store_true", help="skip the front-end / decode roofline measurements")
    ap.add_argument("--no-extra-workloads", action="store_true", help="skip the hubert / wavlm lines and the eager arm")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the torch-eager-on-GPU arm")
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.gpus == 1 or args.impl == "reference":
        protect_stdout()  # (the convenience launcher below leaves stdout to the ranks it spawns)
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus != world and world == 1 and args.gpus > 1:
            # convenience: spawn torchrun ourselves
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", str(Path(__file__).resolve()),
                   "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
                   "--hours", str(args.hours), "--workload", args.workload, "--corpus-files", str(args.corpus_files),
                   "--corpus-model", args.corpus_model, "--corpus-median-s", str(args.corpus_median_s),
                   "--corpus-hours", str(args.corpus_hours)]
            sys.exit(subprocess.call(cmd))
        run_gpu(args)


if __name__ == "__main__":
    main()
