#!/usr/bin/env python
"""Headline benchmark: audio-hours/sec of the sliding-window inference path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (`config.workload`): Whisper-small-dims encoder + segma LSTM/linear heads (`surgical_hydra`),
80-bin log-mel, one 1 h synthetic 16 kHz file per step and per GPU, 4 s windows at the reference's step
(905 forward windows, 179 999 frames), batch 128, default 0.5 thresholds, intervals decoded on the device.
A step = the whole hot path over one file: window -> log-mel -> encoder -> LSTM -> heads -> logits on the
file timeline -> threshold + run-length decode -> interval table (+ NCCL all-gather of tables for N > 1).

  value  : device-timed (CUDA events), PCM already resident in HBM.
  e2e    : the same through the public API (`apply_model_on_audio` + `decode_logits`) from pinned host
           PCM: H2D copy of the file and D2H read of the interval table inside the timed region.
  roofline: the dominant kernel (tcgen05 GEMM / implicit conv), algorithmic FLOPs / CUDA-event time of its
           launches inside the timed region, against the measured sustained bf16 cuBLAS peak.
  cpu_baseline: the oracle (reference arithmetic in torch fp32) on this box's host cores on a bounded
           sample of the same workload.
`--impl reference` times that CPU arm alone (all host threads), same metric / unit / config.
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
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from segma_b200 import ops, synth
    from segma_b200.config import make_config
    from segma_b200.distributed import gather_file_tables, init_from_env
    from segma_b200.encoders import MultiLabelEncoder
    from segma_b200.geometry import INFERENCE_SETTINGS
    from segma_b200.inference import apply_model_on_audio, default_thresholds
    from segma_b200.models import Models
    from segma_b200.thresholds import logit_cut

    # NCCL writes its version banner (NCCL_DEBUG=VERSION and up) to stdout unless told otherwise; stdout carries the
    # single JSON line of this script
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank, world, local = init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ops.device_check()
    n_samples = int(args.hours * HOUR_SAMPLES)
    audio_s = n_samples / 16_000

    le = MultiLabelEncoder(list(LABELS))
    kind, _, gflop_per_window = WORKLOADS[args.workload]
    cfg = make_config(kind)
    if args.workload == "whisper":
        sd = synth.surgical_hydra_state_dict(synth.WHISPER_SMALL, seed=0)
    else:
        sd = synth.hubert_hydra_state_dict(synth.WAVLM_BASE if args.workload == "wavlm" else synth.HUBERT_BASE, seed=0)
    model = Models[kind].from_state_dict(sd, le, cfg).to(dev)
    thr = default_thresholds(le)
    cuts = [logit_cut(0.5)] * len(LABELS)

    host_pcm = torch.from_numpy(synth.synth_audio(n_samples, seed=rank)).pin_memory()
    dev_pcm = host_pcm.to(dev)

    def step_resident():
        logits = apply_model_on_audio(dev_pcm, model, INFERENCE_SETTINGS, dev, batch_size=BATCH)
        table = ops.decode_intervals(logits, cuts, mode=ops.DECODE_LOGIT)
        table[:, 0] = rank
        return gather_file_tables(table) if world > 1 else table

    def step_e2e():
        logits = apply_model_on_audio(host_pcm, model, INFERENCE_SETTINGS, dev, batch_size=BATCH)
        table = ops.decode_intervals(logits, cuts, mode=ops.DECODE_LOGIT)
        table[:, 0] = rank
        full = gather_file_tables(table) if world > 1 else table
        return full.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item() / 1e3, ms[1].item() / 1e3, out

    for _ in range(args.warmup):
        step_resident()
    step_e2e()

    with ClockSampler(local) as clocks:
        ops.stats.reset()
        dev_s, _, table = timed(step_resident, args.steps)
        launches = ops.stats.launches
        _, e2e_wall, table_host = timed(step_e2e, args.steps)
    n_intervals = int(table.shape[0])

    # roofline of the dominant kernel: a separately timed pass with an event pair around every GEMM launch
    ops.stats.reset()
    ops.stats.profile = True
    barrier()
    step_resident()
    torch.cuda.synchronize()
    ops.stats.profile = False
    breakdown = ops.stats.breakdown()
    gemms = [v for k, v in breakdown.items() if k.startswith("segma_gemm_f16")]
    g_ms, g_flops, n_gemm = sum(v["ms"] for v in gemms), sum(v["work"] for v in gemms), sum(v["calls"] for v in gemms)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:  # noqa: BLE001
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    achieved_tf = g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    total_flops_per_step = 905 * gflop_per_window * 1e9 * (n_samples / HOUR_SAMPLES)

    # secondary HBM-bound kernels, timed alone (burst peak): log-mel front end on one 128-window batch and
    # threshold + run-length decode on a >= 100 h batch of logits
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    def time_kernel(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    side = {}
    if rank == 0 and not args.no_side_kernels:
        n_w = 1024  # 1024 windows: 262 MB in + 983 MB out, larger than L2
        span = STEP_SAMPLES * (n_w - 1) + WIN
        pcm_side = dev_pcm[:span] if dev_pcm.numel() >= span else torch.randn(span, device=dev) * 0.1
        out_f32 = torch.empty((n_w, 80, 3000), dtype=torch.float32, device=dev)
        scratch = torch.empty(ops.logmel_scratch_bytes(n_w, WIN), dtype=torch.uint8, device=dev)
        lib = ops._lib()
        st = torch.cuda.current_stream().cuda_stream
        t_mel = time_kernel(lambda: lib.segma_logmel(pcm_side.data_ptr(), pcm_side.numel(), n_w, WIN, STEP_SAMPLES,
                                                     out_f32.data_ptr(), None, scratch.data_ptr(), st))
        mel_bytes = n_w * (4 * WIN + 4 * 80 * 3000)
        side["logmel"] = {"bound": "hbm", "achieved": mel_bytes / t_mel / 1e9, "peak": hbm_peak, "unit": "GB/s",
                          "frac": mel_bytes / t_mel / 1e9 / hbm_peak, "us_per_window": t_mel / n_w * 1e6,
                          "algorithmic_bytes_per_window": 4 * WIN + 4 * 80 * 3000}
        del out_f32, scratch
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

    if rank != 0:
        return
    value = world * (audio_s / 3600.0) * args.steps / dev_s
    e2e_value = world * (audio_s / 3600.0) * args.steps / e2e_wall
    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.workload == "whisper":
        r = cpu_reference_run(args.ref_windows, 1, 0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16",
        "dtype_detail": "f16 tensor-core operands; f32 accumulation, residual stream, LayerNorm/softmax statistics, LSTM state, logits",
        "data": "synthetic", "config": _config(world, audio_s, args.workload),
        "realtime_factor_per_gpu": value * 3600.0 / world,
        "model_tflops_per_gpu": total_flops_per_step * args.steps / dev_s / 1e12,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": host_pcm.numel() * 4,
                "d2h_bytes_per_step": int(table_host.numel()) * 4 + 4},
        "gpu_launches": launches,
        "intervals_per_step": n_intervals,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc5_kernel (tcgen05 GEMM + implicit conv)",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     # dram__bytes_read+write per launch, mean of the four layer GEMMs (QKV, out-proj, fc1, fc2) of a
                     # 128-window batch from profiles/r01e_prof_gemm_ncu_full.txt (1.13 + 1.42 + 1.43 + 2.39 GB) / 4; their
                     # algorithmic bytes: 1.62e9
                     "traffic": 1.59e9 if args.workload == "whisper" else None, "launches": n_gemm, "avg_launch_ms": g_ms / max(n_gemm, 1),
                     "share_of_step": (g_ms / 1e3) / (dev_s / args.steps), "peak_source": peak_src},
        "cpu_baseline": cpu,
        "side_kernels": side,
        "clocks": clocks.summary(),
        "breakdown_ms_per_step": {k: [round(v["ms"], 3), round(v["work"] / max(v["ms"], 1e-9) / 1e9, 1)]
                                  for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="segma_b200", choices=["segma_b200", "reference"])
    ap.add_argument("--hours", type=float, default=1.0, help="audio hours per GPU per step")
    ap.add_argument("--ref-windows", type=int, default=16, help="windows per step of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="whisper", choices=sorted(WORKLOADS),
                    help="whisper = BASELINE config 2 (the headline); hubert / wavlm = configs 1 and 3 models, informational")
    ap.add_argument("--no-side-kernels", action="store_true", help="skip the log-mel / decode roofline measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus != world and world == 1 and args.gpus > 1:
            # convenience: spawn torchrun ourselves
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", str(Path(__file__).resolve()),
                   "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
                   "--hours", str(args.hours)]
            sys.exit(subprocess.call(cmd))
        run_gpu(args)


if __name__ == "__main__":
    main()
