#!/usr/bin/env python
"""Benchmark of the SonicScribe transcription hot path (BASELINE.json metric: RTFx = audio-seconds / second).

    python bench.py --gpus 1 --steps K --warmup W            # this implementation, one B200
    torchrun ... bench.py --gpus N ...                          # one rank (= one model replica) per GPU, weak scaling
    python bench.py --impl reference ...                       # the reference's CPU arithmetic (oracle port) on host cores

A step = one pass of the hot path (peak-norm/PCM16 pre-step + log-mel + encoder + adapter + prefill + greedy KV-cache decode)
over one batch of synthetic 20 s / 16 kHz segments with the full-size GLM-ASR-Nano-2512 geometry and seeded random weights.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEG_SECONDS = 20.0
SEG_SAMPLES = 320000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_segments(batch: int, rank: int):
    from sonicscribe_b200.synth import synth_audio
    return [synth_audio('speech', SEG_SAMPLES, seed=1000 * rank + i) for i in range(batch)]


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's arithmetic (HF GlmAsr graph restated in oracle/) on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(sd, dims, sample_tokens: int, max_new: int, threads: int, dtype_name="bf16"):
    """One 20 s segment through the oracle port: full pre-step + log-mel + encoder + adapter + prefill, then `sample_tokens`
    greedy steps; the per-token cost is extrapolated to `max_new` tokens.  Returns (rtfx, detail)."""
    import torch

    from oracle import mel_oracle as mo
    from oracle import model_oracle as ora

    torch.set_num_threads(threads)
    dt = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    w = {k: v.to(dt) for k, v in sd.items()}
    cfg = ora.OracleConfig(enc_layers=dims.enc_layers, dec_layers=dims.dec_layers)
    x = mo.synth_audio("speech", SEG_SAMPLES, seed=1)
    n_audio = mo.n_audio_tokens(SEG_SAMPLES)
    ids = list(range(100, 108)) + [59260] * n_audio + list(range(200, 212))
    t0 = time.perf_counter()
    mel, _ = mo.log_mel(mo.prestep(x))
    t1 = time.perf_counter()
    new, _, _ = ora.generate_greedy(w, cfg, torch.from_numpy(mel), n_audio, ids, 1)
    t2 = time.perf_counter()
    new, _, _ = ora.generate_greedy(w, cfg, torch.from_numpy(mel), n_audio, ids, 1 + sample_tokens)
    t3 = time.perf_counter()
    t_front = t1 - t0
    t_encprefill = t2 - t1
    t_tok = max((t3 - t2) - t_encprefill, 1e-9) / sample_tokens
    total = t_front + t_encprefill + (max_new - 1) * t_tok
    return SEG_SECONDS / total, {"mel_s": t_front, "enc_prefill_s": t_encprefill, "per_token_s": t_tok, "extrapolated_total_s": total}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from sonicscribe_b200.weights import ModelDims, synthetic_state_dict

    threads = os.cpu_count() or 1
    dims = ModelDims(enc_layers=args.enc_layers, dec_layers=args.dec_layers)
    sd = synthetic_state_dict(dims, seed=0)
    vals = []
    detail = None
    for i in range(args.warmup_ref + args.steps_ref):
        v, detail = cpu_reference_sample(sd, dims, args.ref_sample_tokens, args.max_new, threads)
        if i >= args.warmup_ref:
            vals.append(v)
    value = float(np.mean(vals))
    ms = 1000.0 * SEG_SECONDS / value
    sample = (f"1 segment of 20 s: pre-step + log-mel + encoder + adapter + prefill measured in full, {args.ref_sample_tokens} greedy decode "
              f"steps measured and extrapolated to {args.max_new} tokens; bf16 weights (the reference's CPU dtype, asr.py:61)")
    line = {
        "impl": "reference", "metric": "RTFx (audio-sec/sec), 20 s segments", "value": value, "unit": "audio-seconds/second",
        "n_gpus": args.gpus, "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, args.batch),
        "cpu_baseline": {"value": value, "unit": "audio-seconds/second", "cores": threads, "kind": "port", "sample": sample, "detail": detail},
        "e2e": {"value": value, "unit": "audio-seconds/second", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = /root/reference/backend/asr.py -> transformers GlmAsr generate; its arithmetic restated in oracle/ (torch CPU), all host threads",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch):
    return {
        "workload": f"file transcription (BASELINE.json configs[3] per-GPU share): batches of {batch} VAD-cut 20 s / 16 kHz segments per GPU, "
                    f"reference pre-step + 128-bin log-mel + GLM-ASR-Nano-2512 ({args.enc_layers}+{args.dec_layers} layers, random-init) greedy, "
                    f"max_new_tokens={args.max_new}",
        "segments_per_step_per_gpu": batch, "segment_seconds": SEG_SECONDS, "max_new_tokens": args.max_new,
        "parallelism": f"replicas x{args.gpus} (no collective)", "mode": args.mode,
        "l2": "every decode step streams the 2.9 GB (bf16) weight set, so no step starts with a warm 126 MB L2",
    }


# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this implementation has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from sonicscribe_b200.engine import FLAG_PCM_DEVICE, FLAG_REFERENCE_PRESTEP, Engine, num_audio_tokens
    from sonicscribe_b200.prompt import synthetic_prompt_ids
    from sonicscribe_b200.weights import ModelDims, iter_synthetic_tensors, synthetic_state_dict

    B, G = args.batch, args.max_new
    dims = ModelDims(enc_layers=args.enc_layers, dec_layers=args.dec_layers)
    t0 = time.time()
    eng = Engine(dims.enc_layers, dims.dec_layers, mode=args.mode, device=local, max_batch=B, max_prompt=320, max_new=G)
    # only the rank that times the CPU baseline keeps the 9 GB fp32 checkpoint on the host; everyone else streams it
    need_sd = rank == 0 and world == 1 and not args.no_cpu_baseline
    sd = synthetic_state_dict(dims, seed=0) if need_sd else None
    eng.load_state_dict(sd if need_sd else iter_synthetic_tensors(dims, seed=0))
    if rank == 0:
        log(f"[bench] weights ready in {time.time() - t0:.1f}s; device bytes {eng.device_bytes() / 2**30:.2f} GiB")

    segs = make_segments(B, rank)
    prompts = [synthetic_prompt_ids(num_audio_tokens(SEG_SAMPLES)) for _ in range(B)]
    lens = np.full(B, SEG_SAMPLES, dtype=np.int32)
    offs = (np.arange(B, dtype=np.int64) * SEG_SAMPLES)
    host = torch.from_numpy(np.concatenate(segs)).pin_memory()           # pinned host PCM for the e2e leg
    dev = host.cuda(non_blocking=False)                                  # HBM-resident PCM for the device leg
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return eng.transcribe_packed(dev.data_ptr(), offs, lens, prompts, G, FLAG_REFERENCE_PRESTEP | FLAG_PCM_DEVICE)

    def step_e2e():
        return eng.transcribe_packed(host.data_ptr(), offs, lens, prompts, G, FLAG_REFERENCE_PRESTEP)

    def timed(fn, steps):
        barrier()
        eng.timer_begin()
        w0 = time.perf_counter()
        out = None
        for _ in range(steps):
            out = fn()
        ms = eng.timer_end()
        wall = (time.perf_counter() - w0) * 1000.0
        barrier()
        t = torch.tensor([ms, wall], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), out

    for _ in range(max(args.warmup, 3)):
        step_device()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, dev_wall, out = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - launches0
    stage = eng.stage_times()
    step_e2e()
    e2e_ms, e2e_wall, out2 = timed(step_e2e, args.steps)
    assert out == out2, "device-resident and host-input passes disagree"

    # p50 latency of ONE 20 s segment alone (the reference's batch-1 call, BASELINE.json metric second half)
    lat = []
    for _ in range(5):
        t0 = time.perf_counter()
        eng.transcribe_packed(host.data_ptr(), offs[:1], lens[:1], prompts[:1], G, FLAG_REFERENCE_PRESTEP)
        lat.append((time.perf_counter() - t0) * 1000.0)
    lat_p50 = float(np.median(lat[1:]))

    # one extra, eager, event-bracketed step: device time per launch class (basis of the roofline object)
    eng.profile_begin()
    step_device()
    prof = eng.profile_end()

    audio_s = world * B * SEG_SECONDS * args.steps
    value = audio_s / (dev_ms / 1000.0)
    e2e_value = audio_s / (e2e_wall / 1000.0)

    if rank == 0:
        peaks = measured_peaks()
        hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        total_prof = sum(v["ms"] for v in prof.values())
        pc = prof.get("dec_persistent_step", {"ms": 0.0, "launches": 0})
        if pc["launches"] > 0:
            # dominant kernel: the persistent decode step (one launch per generated token for the whole batch)
            S = len(prompts[0])
            w_bytes = 1.472e9 * (2 if args.mode != "fp32" else 4)          # every decoder + lm_head weight once per step
            kv_bytes = B * 57344.0 * (S + G / 2.0)                          # K and V of 28 layers over the mean context
            bytes_per_launch = w_bytes + kv_bytes
            c = pc
            kname = ("decode_persist_kernel: cooperative per-token kernel (28 layers + lm_head + greedy pick; weight streaming by "
                     + ("TMA + tcgen05 with split-K partials" if 32 < B <= 64 and args.mode != "int8" else "mma.sync with CTA-level split-K")
                     + ", fused RoPE/KV-append/attention/RMSNorm/SwiGLU)")
        else:
            esz = 2
            bytes_per_launch = 2 * 6144 * 2048 * esz + B * 2048 * esz + B * 6144 * esz      # weights + activations in/out
            c = prof["dec_gemm_gateup"]
            kname = "gemm_tc_kernel<swap> gate/up projection of the greedy decode step (weight streaming, SwiGLU epilogue)"
        avg_ms = c["ms"] / max(c["launches"], 1)
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (profiles/r01_traffic.json, written
        # by scripts/ncu_traffic.py from dram__bytes_read.sum + dram__bytes_write.sum); only quoted for the batch it was taken at
        traffic = None
        try:
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_traffic.json")))
            if pc["launches"] > 0 and tj.get("kernel") == "decode_persist_kernel" and tj.get("batch") == B and tj.get("mode") == args.mode:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
        roofline = {
            "kernel": kname, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
            "launches_timed": c["launches"], "share_of_step": c["ms"] / total_prof if total_prof > 0 else None,
        }
        cpu_base = None
        if not args.no_cpu_baseline and world == 1:          # the CPU baseline is timed on rank 0 of the 1-GPU run only
            t1 = time.time()
            v, detail = cpu_reference_sample(sd, dims, args.ref_sample_tokens, G, os.cpu_count() or 1)
            cpu_base = {"value": v, "unit": "audio-seconds/second", "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"1 segment of 20 s, front end + encoder + prefill in full, {args.ref_sample_tokens} decode steps extrapolated "
                                  f"to {G}; oracle port (torch CPU bf16) of the HF graph the reference runs", "detail": detail,
                        "seconds_spent": time.time() - t1}
        line = {
            "metric": "RTFx (audio-sec/sec), 20 s segments", "value": value, "unit": "audio-seconds/second", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.mode != "fp32" else "f32", "data": "synthetic",
            "config": workload_config(args, B),
            "e2e": {"value": e2e_value, "unit": "audio-seconds/second", "h2d_bytes_per_step": int(B * SEG_SAMPLES * 4 + sum(len(p) for p in prompts) * 4 * 4),
                    "d2h_bytes_per_step": int(B * G * 4 + B * 4), "ms_per_step_wall": e2e_wall / args.steps,
                    "api": "Engine.transcribe_packed -> sonic_transcribe_batch (host PCM, pinned)"},
            "p50_latency_ms_single_20s_segment": lat_p50, "latency_ms_per_batch": dev_ms / args.steps,
            "gpu_launches": int(launches), "stage_ms_last_step": stage, "wall_ms_per_step": dev_wall / args.steps,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_base,
            "profile_ms_by_class": {k: round(v["ms"], 3) for k, v in prof.items()},
            "profile_launches_by_class": {k: v["launches"] for k, v in prof.items()},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("SONIC_BENCH_BATCH", "64")))
    ap.add_argument("--max-new", type=int, default=128)
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32", "int8"])
    ap.add_argument("--enc-layers", type=int, default=32)
    ap.add_argument("--dec-layers", type=int, default=28)
    ap.add_argument("--ref-sample-tokens", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # the reference arm is CPU-bound (~8 s per sample on 8 cores): bound its repetitions so the run ends within minutes
    args.steps_ref = max(1, min(args.steps, 3))
    args.warmup_ref = 1 if args.warmup > 0 else 0
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
