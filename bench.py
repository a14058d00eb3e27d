#!/usr/bin/env python
"""Benchmark of the SonicScribe transcription hot path (BASELINE.json metric: RTFx = audio-seconds / second).

    python bench.py --gpus 1 --steps K --warmup W            # this implementation, one B200 (default workload)
    torchrun ... bench.py --gpus N ...                          # one rank (= one model replica) per GPU
    python bench.py --impl reference ...                       # the reference's own CPU arithmetic (HF classes) on host cores
    python bench.py --workload realtime                        # BASELINE config 2: interim / committed call latency
    python bench.py --workload file1h   [under torchrun]       # BASELINE config 4: 1 h = 180 x 20 s, sharded, in order
    python bench.py --mode int8 --batch 1                      # BASELINE config 3: INT8 batch-1 decode
    (BASELINE config 5, the log-mel sweep: scripts/bench_mel.py)

Default workload: a step = one pass of the hot path (peak-norm/PCM16 pre-step + log-mel + encoder + adapter + prefill +
greedy KV-cache decode) over one batch of synthetic 20 s / 16 kHz segments with the full-size GLM-ASR-Nano-2512 geometry and
seeded random weights.
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
N_PARAMS_DEC_LAYERS = 1.3506e9      # 28 Llama layers
N_PARAMS_LM_HEAD = 0.1214e9         # untied lm_head (stays bf16 in int8 mode)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def decode_weight_bytes(mode: str) -> float:
    """Algorithmic weight bytes one greedy step streams: bf16 2.944 GB, int8 1.593 GB (+ row scales), fp32 5.888 GB."""
    if mode == "int8":
        return N_PARAMS_DEC_LAYERS * 1 + N_PARAMS_LM_HEAD * 2 + 28 * (3072 + 2048 + 12288 + 2048) * 4
    return (N_PARAMS_DEC_LAYERS + N_PARAMS_LM_HEAD) * (4 if mode == "fp32" else 2)


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
# CPU arms.  (1) --impl reference: the classes /root/reference/backend/asr.py calls — transformers' WhisperFeatureExtractor
# and GlmAsrForConditionalGeneration.generate(do_sample=False) — run unmodified on the host cores in the reference's CPU
# dtype (bf16, asr.py:61,129-130), every one of the max_new_tokens timed.  asr.py's own wrapper cannot be imported on this
# image (soundfile / torchaudio absent), so its pre-step (peak-normalise + PCM_16 round trip) comes from oracle/mel_oracle.py.
# (2) cpu_baseline of the GPU line: the oracle port (oracle/model_oracle.py), a bounded sample.
# ----------------------------------------------------------------------------------------------------------------------
def build_hf_reference(dims, sd, dtype):
    import torch
    from transformers import GlmAsrConfig, GlmAsrForConditionalGeneration

    cfg = GlmAsrConfig(audio_config={"num_hidden_layers": dims.enc_layers}, text_config={"num_hidden_layers": dims.dec_layers})
    torch.manual_seed(0)
    model = GlmAsrForConditionalGeneration(cfg).to(dtype)
    model.load_state_dict({k: v.to(dtype) for k, v in sd.items()}, strict=True, assign=True)
    model.generation_config.pad_token_id = 59246
    return model.eval()


def hf_reference_step(model, fe, x, max_new, dtype):
    """One transcribe() worth of work on one 20 s segment, as asr.py:393-422 orders it."""
    import torch

    from oracle import mel_oracle as mo

    t0 = time.perf_counter()
    xp = mo.prestep(x)                                                            # asr.py:248-276
    f = fe(xp, sampling_rate=16000, return_attention_mask=True, padding="max_length", return_tensors="pt")
    n_audio = mo.n_audio_tokens(x.shape[0])
    ids = list(range(100, 108)) + [59260] * n_audio + list(range(200, 212))
    kw = dict(input_ids=torch.tensor([ids]), input_features=f["input_features"].to(dtype), input_features_mask=f["attention_mask"],
              attention_mask=torch.ones(1, len(ids), dtype=torch.long))
    with torch.no_grad():
        seq = model.generate(**kw, max_new_tokens=max_new, do_sample=False)
    new = seq[0, len(ids):].tolist()
    return time.perf_counter() - t0, new


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from transformers import WhisperFeatureExtractor

    from sonicscribe_b200.synth import synth_audio
    from sonicscribe_b200.weights import ModelDims, synthetic_state_dict

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    dims = ModelDims(enc_layers=args.enc_layers, dec_layers=args.dec_layers)
    dtype = torch.bfloat16
    t0 = time.time()
    model = build_hf_reference(dims, synthetic_state_dict(dims, seed=0), dtype)
    fe = WhisperFeatureExtractor(feature_size=128)
    log(f"[bench/reference] HF model ready in {time.time() - t0:.1f}s on {threads} threads")
    budget_s = float(os.environ.get("SONIC_REF_BUDGET_S", "240"))
    times, n_tok = [], 0
    warm = 1 if args.warmup > 0 else 0
    for i in range(warm):
        hf_reference_step(model, fe, synth_audio("speech", SEG_SAMPLES, seed=1), args.max_new, dtype)
    t_start = time.time()
    for i in range(args.steps):
        dt, new = hf_reference_step(model, fe, synth_audio("speech", SEG_SAMPLES, seed=1000 + i), args.max_new, dtype)
        times.append(dt)
        n_tok = len(new)
        if time.time() - t_start + dt > budget_s and len(times) >= 2:
            break
    per_step = float(np.mean(times))
    value = SEG_SECONDS / per_step
    sample = (f"{len(times)} timed steps of ONE 20 s segment each (the reference serves batch 1: transcription_manager.py:53-54, main.py:610-612), "
              f"pre-step + WhisperFeatureExtractor + GlmAsrForConditionalGeneration.generate(do_sample=False), all {n_tok} of "
              f"max_new_tokens={args.max_new} steps timed, bf16 (the reference's CPU dtype); steps capped by a {budget_s:.0f} s budget")
    line = {
        "impl": "reference", "metric": "RTFx (audio-sec/sec), 20 s segments", "value": value, "unit": "audio-seconds/second",
        "n_gpus": args.gpus, "steps": len(times), "warmup": warm, "ms_per_step": per_step * 1000.0, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, args.batch),
        "cpu_baseline": {"value": value, "unit": "audio-seconds/second", "cores": threads, "kind": "reference", "sample": sample,
                         "step_seconds": [round(t, 3) for t in times]},
        "e2e": {"value": value, "unit": "audio-seconds/second", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = /root/reference/backend/asr.py -> transformers 5.5.0 WhisperFeatureExtractor + GlmAsrForConditionalGeneration.generate, "
                "imported and run as shipped on all host threads; asr.py's wrapper itself needs soundfile/torchaudio (absent), its pre-step is "
                "restated by oracle/mel_oracle.py",
    }
    print(json.dumps(line), flush=True)


def cpu_port_sample(sd, dims, sample_tokens: int, max_new: int, threads: int, dtype_name="bf16"):
    """cpu_baseline of the GPU line: one 20 s segment through the oracle port: full pre-step + log-mel + encoder + adapter +
    prefill, then `sample_tokens` greedy steps; the per-token cost is extrapolated to `max_new` tokens."""
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
    # The stages are timed directly (the loop of oracle.generate_greedy restated around the oracle's own stage functions): an
    # earlier version took the per-token cost as the DIFFERENCE of two whole passes, which a cold first pass once made 20 x
    # too small.  One untimed front end + encoder pass first (weights paged in, thread pool started).
    mel, _ = mo.log_mel(mo.prestep(x))
    ora.encoder_forward(w, cfg, torch.from_numpy(mel))
    t0 = time.perf_counter()
    mel, _ = mo.log_mel(mo.prestep(x))
    t1 = time.perf_counter()
    with torch.no_grad():
        tids = torch.as_tensor(ids, dtype=torch.long)
        enc = ora.encoder_forward(w, cfg, torch.from_numpy(mel))
        ae = ora.adapter_forward(w, cfg, enc, n_audio)
        xs = ora.embed_merge(w, tids, ae)
        cache = ora.KVCache(cfg.dec_layers)
        logits = ora.decoder_forward(w, cfg, xs, 0, cache)
        t2 = time.perf_counter()
        pos = tids.shape[0]
        for _ in range(sample_tokens):
            tok = int(torch.argmax(logits.float()))
            xs = w["language_model.model.embed_tokens.weight"][tok][None]
            logits = ora.decoder_forward(w, cfg, xs, pos, cache)
            pos += 1
        t3 = time.perf_counter()
    t_front = t1 - t0
    t_encprefill = t2 - t1
    t_tok = (t3 - t2) / sample_tokens
    total = t_front + t_encprefill + (max_new - 1) * t_tok
    return SEG_SECONDS / total, {"mel_s": t_front, "enc_prefill_s": t_encprefill, "per_token_s": t_tok, "extrapolated_total_s": total}


def workload_config(args, batch):
    return {
        "workload": f"file transcription (BASELINE.json configs[3] per-GPU share): batches of {batch} VAD-cut 20 s / 16 kHz segments per GPU, "
                    f"reference pre-step + 128-bin log-mel + GLM-ASR-Nano-2512 ({args.enc_layers}+{args.dec_layers} layers, random-init) greedy, "
                    f"max_new_tokens={args.max_new}",
        "segments_per_step_per_gpu": batch, "segment_seconds": SEG_SECONDS, "max_new_tokens": args.max_new,
        "parallelism": f"replicas x{args.gpus} (no collective)", "mode": args.mode,
        "l2": "every decode step streams the whole decoder weight set (2.9 GB bf16 / 1.6 GB int8), so no step starts with a warm 126 MB L2",
    }


def dist_setup():
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
    return world, rank, local


def make_engine(args, local, rank, world, max_batch, max_prompt=320, keep_sd=False):
    from sonicscribe_b200.engine import Engine
    from sonicscribe_b200.weights import ModelDims, iter_synthetic_tensors, synthetic_state_dict

    dims = ModelDims(enc_layers=args.enc_layers, dec_layers=args.dec_layers)
    t0 = time.time()
    eng = Engine(dims.enc_layers, dims.dec_layers, mode=args.mode, device=local, max_batch=max_batch, max_prompt=max_prompt, max_new=max(args.max_new, 16))
    sd = synthetic_state_dict(dims, seed=0) if keep_sd else None
    eng.load_state_dict(sd if keep_sd else iter_synthetic_tensors(dims, seed=0))
    if rank == 0:
        log(f"[bench] weights ready in {time.time() - t0:.1f}s; device bytes {eng.device_bytes() / 2**30:.2f} GiB")
    return eng, dims, sd


# ----------------------------------------------------------------------------------------------------------------------
# default workload: batches of 20 s segments
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world, rank, local = dist_setup()
    from sonicscribe_b200.asr import ASRModel
    from sonicscribe_b200.engine import FLAG_PCM_DEVICE, FLAG_REFERENCE_PRESTEP, num_audio_tokens
    from sonicscribe_b200.prompt import synthetic_prompt_ids

    B, G = args.batch, args.max_new
    # only the rank that times the CPU baseline keeps the 9 GB fp32 checkpoint on the host; everyone else streams it
    need_sd = rank == 0 and world == 1 and not args.no_cpu_baseline
    eng, dims, sd = make_engine(args, local, rank, world, B, keep_sd=need_sd)

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

    min_warm = int(os.environ.get('SONIC_BENCH_MINWARM', '3'))      # 3 for any reported number; profiling runs may lower it
    for _ in range(max(args.warmup, min_warm)):
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

    # the UNCHANGED interface under the reference's own concurrency: N host threads inside ASRModel.transcribe on one instance
    # (main.py:429-445 uses 3 executor threads); the dynamic batcher coalesces them into one device pass
    api_threads = {}
    if rank == 0 and not args.no_api_threads:
        asr = ASRModel("synthetic", device=f"cuda:{local}", mode={"bf16": "native"}.get(args.mode, args.mode), engine=eng)
        seg_t = [torch.from_numpy(s)[None] for s in segs]
        for nt in (1, 3, 16):
            if nt > B:
                continue
            reps = 2 if nt < 16 else 3

            def work(i):
                for r in range(reps):
                    asr.transcribe(seg_t[(i + r * nt) % B], max_new_tokens=G)

            work(0)
            s0 = asr.batcher_stats()
            t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
            [t.start() for t in th]; [t.join() for t in th]
            wall = time.perf_counter() - t0
            s1 = asr.batcher_stats()
            api_threads[str(nt)] = {"rtfx": nt * reps * SEG_SECONDS / wall, "calls": nt * reps, "wall_s": wall,
                                    "device_passes": s1["batches"] - s0["batches"]}
        asr.close()

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
        S = len(prompts[0])
        if pc["launches"] > 0:
            # dominant kernel: the persistent decode step (one launch per generated token for the whole batch)
            w_bytes = decode_weight_bytes(args.mode)                         # every decoder + lm_head weight once per step
            kv_bytes = B * 57344.0 * (S + G / 2.0)                          # K and V of 28 layers over the mean context
            bytes_per_launch = w_bytes + kv_bytes
            c = pc
            kname = "decode_persist: cooperative per-token kernel (28 layers + lm_head + greedy pick; weight streaming by TMA + tcgen05, " \
                    "fused RoPE/KV-append/attention/RMSNorm/SwiGLU)"
        else:
            esz = 2
            bytes_per_launch = 2 * 6144 * 2048 * esz + B * 2048 * esz + B * 6144 * esz      # weights + activations in/out
            c = prof["dec_gemm_gateup"]
            kname = "gemm_tc_kernel<swap> gate/up projection of the greedy decode step (weight streaming, SwiGLU epilogue)"
        avg_ms = c["ms"] / max(c["launches"], 1)
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (profiles/*_traffic.json, written
        # by scripts/ncu_traffic.py from dram__bytes_read.sum + dram__bytes_write.sum); only quoted for the batch it was taken at
        traffic = None
        for tf in ("r02_traffic.json", "r01_traffic.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", tf)))
                if pc["launches"] > 0 and tj.get("batch") == B and tj.get("mode") == args.mode:
                    traffic = tj.get("dram_bytes_per_launch")
                    break
            except Exception:
                continue
        roofline = {
            "kernel": kname, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
            "launches_timed": c["launches"], "share_of_step": c["ms"] / total_prof if total_prof > 0 else None,
        }
        cpu_base = None
        if not args.no_cpu_baseline and world == 1:          # the CPU baseline is timed on rank 0 of the 1-GPU run only
            t1 = time.time()
            v, detail = cpu_port_sample(sd, dims, args.ref_sample_tokens, G, os.cpu_count() or 1)
            cpu_base = {"value": v, "unit": "audio-seconds/second", "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"1 segment of 20 s, front end + encoder + prefill in full, {args.ref_sample_tokens} decode steps extrapolated "
                                  f"to {G}; oracle port (torch CPU bf16) of the HF graph the reference runs", "detail": detail,
                        "seconds_spent": time.time() - t1}
        dec_ms = stage.get("decode_ms", 0.0)
        line = {
            "metric": "RTFx (audio-sec/sec), 20 s segments", "value": value, "unit": "audio-seconds/second", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.mode != "fp32" else "f32", "data": "synthetic",
            "config": workload_config(args, B),
            "e2e": {"value": e2e_value, "unit": "audio-seconds/second", "h2d_bytes_per_step": int(B * SEG_SAMPLES * 4 + sum(len(p) for p in prompts) * 4 * 4),
                    "d2h_bytes_per_step": int(B * G * 4 + B * 4), "ms_per_step_wall": e2e_wall / args.steps,
                    "api": "Engine.transcribe_packed -> sonic_transcribe_batch (host PCM, pinned)"},
            "p50_latency_ms_single_20s_segment": lat_p50, "latency_ms_per_batch": dev_ms / args.steps,
            "api_threads_asrmodel_transcribe": api_threads,
            "decode": {"ms_per_token_step": dec_ms / max(G - 1, 1), "tokens_per_s": B * (G - 1) / (dec_ms / 1000.0) if dec_ms > 0 else None,
                       "weight_bytes_per_step": decode_weight_bytes(args.mode),
                       "weight_gbps": decode_weight_bytes(args.mode) * (G - 1) / (dec_ms / 1000.0) / 1e9 if dec_ms > 0 else None},
            "gpu_launches": int(launches), "stage_ms_last_step": stage, "wall_ms_per_step": dev_wall / args.steps,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_base,
            "profile_ms_by_class": {k: round(v["ms"], 3) for k, v in prof.items()},
            "profile_launches_by_class": {k: v["launches"] for k, v in prof.items()},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE config 2: single-stream realtime.  64 ms chunks arrive as int16; while the speaker talks an interim result is
# requested at most once per second on the last 20 chunks (20 480 samples, 15 tokens: connection_manager.py:88-92,127-166,
# audio_manager.py:106-114, transcription_manager.py:25); at the end of the utterance the whole <= 20 s segment is decoded
# with min(50 + 5*dur, 200) = 150 tokens (transcription_manager.py:37).  Latency = wall clock of the call, int16 bytes in,
# text out, through TranscriptionManager -> ASRModel.transcribe_pcm16.
# ----------------------------------------------------------------------------------------------------------------------
def run_realtime(args):
    import asyncio

    import torch

    world, rank, local = dist_setup()
    if rank != 0:
        return
    import sonicscribe_b200.models_manager as mm
    from sonicscribe_b200.asr import ASRModel
    from sonicscribe_b200.synth import synth_audio
    from sonicscribe_b200.transcription_manager import TranscriptionManager

    eng, dims, _ = make_engine(args, local, rank, world, max_batch=max(args.batch if args.batch <= 16 else 4, 1), max_prompt=448)
    asr = ASRModel("synthetic", device=f"cuda:{local}", mode={"bf16": "native"}.get(args.mode, args.mode), engine=eng)
    mm._asr_model = asr
    mgr = TranscriptionManager()
    utt = synth_audio("speech", SEG_SAMPLES, seed=7)
    pcm = np.clip(np.rint(utt * 32767.0), -32768, 32767).astype(np.int16)
    n_interim = 19                                   # one per second of a 20 s utterance
    loop = asyncio.new_event_loop()

    def interim(k):      # after second k+1: the last 20 chunks of 1024 samples
        end = (k + 1) * 16000 // 1024 * 1024
        chunk = pcm[max(0, end - 20 * 1024):end].tobytes()
        t0 = time.perf_counter()
        text = loop.run_until_complete(mgr.transcribe_temporary(chunk))
        return (time.perf_counter() - t0) * 1000.0, text

    def committed():
        t0 = time.perf_counter()
        text = loop.run_until_complete(mgr.transcribe_committed(pcm.tobytes(), SEG_SECONDS))
        return (time.perf_counter() - t0) * 1000.0, text

    for _ in range(max(args.warmup, 3)):
        interim(3); committed()
    sampler = ClockSampler(local)
    sampler.start()
    lat_i, lat_c, stage_i, stage_c = [], [], None, None
    launches0 = eng.launch_count()
    t_all = time.perf_counter()
    for u in range(args.steps):                      # `steps` utterances of 20 s each
        for k in range(n_interim):
            ms, text = interim(k)
            assert 1 <= len(text.split()) <= 15
            lat_i.append(ms)
        stage_i = eng.stage_times()
        ms, text = committed()
        assert 1 <= len(text.split()) <= 150
        lat_c.append(ms)
        stage_c = eng.stage_times()
    wall = time.perf_counter() - t_all
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    p = lambda a, q: float(np.percentile(np.array(a), q))
    peaks = measured_peaks()
    hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)
    dec_ms = stage_c["decode_ms"] / 149.0
    bytes_step = decode_weight_bytes(args.mode) + 57344.0 * (270 + 75)
    line = {
        "metric": "realtime single stream: p50 latency of the interim call (1.28 s window, 15 tokens), ms", "value": p(lat_i, 50), "unit": "ms",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": wall * 1000.0 / args.steps, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.mode != "fp32" else "f32", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: one realtime stream, int16 64 ms chunks; interim decode once per second on the last 20 chunks "
                               "(20480 samples, max_new_tokens=15), committed decode of the 20 s utterance (max_new_tokens=150); full-size random-init "
                               "GLM-ASR-Nano-2512; latency = wall clock of TranscriptionManager.transcribe_temporary/committed (bytes in, text out)",
                   "utterances": args.steps, "interim_calls": len(lat_i), "mode": args.mode,
                   "interim_encoder_window": ("short (opt-in SONIC_FLAG_SHORT_WINDOW via SONIC_SHORT_WINDOW_MAX_NEW: NOT the reference's numerics)"
                                              if int(os.environ.get("SONIC_SHORT_WINDOW_MAX_NEW", "0")) >= 15 else "full 30 s window (the reference's)")},
        "interim_ms": {"p50": p(lat_i, 50), "p95": p(lat_i, 95), "max": max(lat_i), "stage_ms": stage_i},
        "committed_ms": {"p50": p(lat_c, 50), "p95": p(lat_c, 95), "max": max(lat_c), "stage_ms": stage_c},
        "interim_budget_ms": 1000.0, "interim_share_of_budget": p(lat_i, 95) / 1000.0,
        "stream_duty_cycle": (sum(lat_i) + sum(lat_c)) / 1000.0 / (args.steps * SEG_SECONDS),
        "e2e": {"value": p(lat_i, 50), "unit": "ms", "h2d_bytes_per_step": 20480 * 2, "d2h_bytes_per_step": 15 * 4 + 4,
                "api": "TranscriptionManager.transcribe_temporary -> ASRModel.transcribe_pcm16 -> sonic_transcribe_batch (int16 host bytes)"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "decode_persist (batch 1) inside the committed call", "bound": "hbm", "achieved": bytes_step / (dec_ms * 1e-3) / 1e9,
                     "peak": hbm_peak, "unit": "GB/s", "frac": bytes_step / (dec_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                     "avg_launch_ms": dec_ms, "bytes_per_launch": bytes_step},
        "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)
    mm._asr_model = None
    asr.close()
    eng.close()


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE config 4: one hour of audio -> cut_long_segments(max 20 s) = 180 segments (main.py:276-287,527-567 with VAD off)
# -> sharded over the ranks (segment i -> rank i mod W) -> each rank: ASRModel.transcribe_batch (dynamic batcher, up to
# `batch` segments per device pass) -> results gathered and re-ordered by segment index (main.py:448-468 streams in order).
# STRONG scaling: the total work is fixed.  Timed region per rank: host PCM -> ids of its shard; max over ranks.
# ----------------------------------------------------------------------------------------------------------------------
def run_file1h(args):
    import torch
    import torch.distributed as dist

    world, rank, local = dist_setup()
    from sonicscribe_b200.asr import ASRModel
    from sonicscribe_b200.pool import cut_long_segments, gather_in_order, shard_indices
    from sonicscribe_b200.synth import synth_audio

    total_samples = int(3600 * 16000)
    cuts = cut_long_segments(0, total_samples, 16000, SEG_SECONDS)
    n_seg = len(cuts)
    mine = shard_indices(n_seg, world, rank)
    B = min(args.batch, 256)
    eng, dims, _ = make_engine(args, local, rank, world, max_batch=B)
    asr = ASRModel("synthetic", device=f"cuda:{local}", mode={"bf16": "native"}.get(args.mode, args.mode), engine=eng)
    # the hour is the concatenation of per-cut synthetic utterances (seed = cut index): every rank materialises only its shard
    segs = [torch.from_numpy(synth_audio("speech", e - s, seed=5000 + i))[None].pin_memory() for i, (s, e) in ((i, cuts[i]) for i in mine)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass():
        return asr.transcribe_ids(segs, max_new_tokens=args.max_new)

    for _ in range(max(min(args.warmup, 2), 1)):
        one_pass()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    walls = []
    launches0 = eng.launch_count()
    res = None
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        res = one_pass()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        barrier()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        walls.append(float(t[0]))
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - launches0
    ordered = gather_in_order([(i, r[:4], len(r)) for i, r in zip(mine, res)], n_seg, world, rank)
    if rank == 0:
        assert [o[0] for o in ordered] == list(range(n_seg)), "results are not in segment order"
        assert all(o[2] == args.max_new for o in ordered)
        wall = float(np.median(walls))
        st = asr.batcher_stats()
        line = {
            "metric": "RTFx (audio-sec/sec), 1 h file = 180 x 20 s segments, strong scaling", "value": 3600.0 / wall, "unit": "audio-seconds/second",
            "n_gpus": world, "steps": args.steps, "warmup": max(min(args.warmup, 2), 1), "ms_per_step": wall * 1000.0, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.mode != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[3]: 1 h of 16 kHz audio cut into {n_seg} segments of 20 s (cut_long_segments, main.py:527-567), "
                                   f"segment i on rank i mod {world}, up to {B} segments per device pass, max_new_tokens={args.max_new}, results gathered in "
                                   "segment order; full-size random-init GLM-ASR-Nano-2512",
                       "segments": n_seg, "segments_per_rank": len(mine), "mode": args.mode, "parallelism": f"replicas x{world} (no data-path collective)"},
            "e2e": {"value": 3600.0 / wall, "unit": "audio-seconds/second", "h2d_bytes_per_step": int(sum(s.numel() for s in segs) * 4),
                    "d2h_bytes_per_step": int(len(segs) * (args.max_new + 1) * 4),
                    "api": "ASRModel.transcribe_ids (host float32 tensors) -> dynamic batcher -> sonic_transcribe_batch"},
            "pass_seconds": [round(w, 4) for w in walls], "batcher": st, "gpu_launches": int(launches), "clocks": clocks,
            "first_ids_of_segment_0_and_last": [ordered[0][1], ordered[-1][1]], "roofline": None, "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    asr.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="file20s", choices=["file20s", "realtime", "file1h"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("SONIC_BENCH_BATCH", "256")))
    ap.add_argument("--max-new", type=int, default=128)
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32", "int8"])
    ap.add_argument("--enc-layers", type=int, default=32)
    ap.add_argument("--dec-layers", type=int, default=28)
    ap.add_argument("--ref-sample-tokens", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-api-threads", action="store_true")
    args = ap.parse_args()
    if args.mode == "int8" and args.batch > 128:
        args.batch = 128                       # the int8 tcgen05 decode class tiles at most 128 segments per launch
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "realtime":
        args.max_new = max(args.max_new, 200)
        run_realtime(args)
    elif args.workload == "file1h":
        run_file1h(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
