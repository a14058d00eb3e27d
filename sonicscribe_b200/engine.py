"""ctypes binding of ``libsonic_b200.so`` (C ABI: ``include/sonic_b200.h``) and a thin numpy-facing ``Engine``.

There is NO CPU fallback: if the shared library is missing or no sm_100a device is present, construction raises.
Replaces, on the device, the arithmetic the reference reaches through ``transformers``
(/root/reference/backend/asr.py:393-422).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

_LIB_NAME = "libsonic_b200.so"
_lib = None

MODE_BF16, MODE_FP32, MODE_INT8 = 0, 1, 2
MODES = {"bf16": MODE_BF16, "native": MODE_BF16, "fp32": MODE_FP32, "int8": MODE_INT8}
FLAG_PEAK_NORM, FLAG_PCM16, FLAG_PCM_S16, FLAG_PCM_DEVICE, FLAG_OUT_DEVICE, FLAG_FEATURES_ONLY = 0x01, 0x02, 0x04, 0x10, 0x20, 0x40
FLAG_SHORT_WINDOW = 0x80      # opt-in streaming encoder for interim calls (include/sonic_b200.h)
ERR_UNKNOWN_TENSOR = -2
FLAG_REFERENCE_PRESTEP = FLAG_PEAK_NORM | FLAG_PCM16

N_MELS, N_FRAMES, MERGED, DEC_HIDDEN, VOCAB = 128, 3000, 375, 2048, 59264
AUDIO_TOKEN_ID = 59260


class SonicConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("device", "mode", "enc_layers", "dec_layers", "max_batch", "max_prompt", "max_new", "debug")]


def lib_path() -> str:
    # SONIC_LIB points at an alternative build of the same library (A/B measurements); default is the in-tree build
    return os.environ.get("SONIC_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load_library():
    """dlopen the in-tree shared library and declare every prototype of include/sonic_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"or sonicscribe_b200/csrc/build.sh (there is no CPU fallback)")
    lib = C.CDLL(path)
    H = C.c_void_p
    i32p, i64p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float)
    protos = {
        "sonic_create": (C.c_int, [C.POINTER(SonicConfig), C.POINTER(H)]),
        "sonic_destroy": (C.c_int, [H]),
        "sonic_last_error": (C.c_char_p, [H]),
        "sonic_version": (C.c_char_p, []),
        "sonic_load_tensor": (C.c_int, [H, C.c_char_p, C.c_void_p, C.c_int32, i64p, C.c_int32]),
        "sonic_finalize_weights": (C.c_int, [H]),
        "sonic_mel": (C.c_int, [H, C.c_void_p, i64p, i32p, C.c_int32, C.c_int32, C.c_void_p, i32p]),
        "sonic_encode": (C.c_int, [H, C.c_int32, f32p, i32p]),
        "sonic_generate": (C.c_int, [H, i32p, i32p, C.c_int32, C.c_int32, i32p, i32p, f32p]),
        "sonic_transcribe_batch": (C.c_int, [H, C.c_void_p, i64p, i32p, C.c_int32, C.c_int32, i32p, i32p, C.c_int32, i32p, i32p, f32p]),
        "sonic_num_audio_tokens": (C.c_int32, [C.c_int64]),
        "sonic_sync": (C.c_int, [H]),
        "sonic_timer_begin": (C.c_int, [H]),
        "sonic_timer_end": (C.c_int, [H, f32p]),
        "sonic_stage_times": (C.c_int, [H, f32p]),
        "sonic_launch_count": (C.c_int64, [H]),
        "sonic_device_bytes": (C.c_int64, [H]),
        "sonic_profile_begin": (C.c_int, [H]),
        "sonic_profile_end": (C.c_int, [H, f32p, i64p, C.c_int32]),
        "sonic_profile_num_classes": (C.c_int32, []),
        "sonic_profile_class_name": (C.c_char_p, [C.c_int32]),
        "sonic_debug_read": (C.c_int, [H, C.c_char_p, f32p, C.c_size_t, C.POINTER(C.c_size_t)]),
        "sonic_debug_set_logit_steps": (C.c_int, [H, i32p, C.c_int32]),
        "sonic_test_enc_attention": (C.c_int, [H, C.c_int32, f32p, f32p, C.c_int32, C.c_int32]),
        "sonic_test_gemm_int8": (C.c_int, [H, C.c_int32, f32p, f32p, f32p, f32p, f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
        "sonic_bench_gemm": (C.c_int, [H, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, f32p]),
        "sonic_bench_mma": (C.c_int, [H, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, f32p, f32p]),
        "sonic_test_gemm": (C.c_int, [H, C.c_int32, C.c_int32, f32p, f32p, f32p, f32p, f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    lib._sonic_protos = tuple(protos)
    _lib = lib
    return lib


def num_audio_tokens(n_samples: int) -> int:
    """transformers/models/glmasr/processing_glmasr.py:97-103 on the frame mask of an n-sample segment."""
    f = -(-min(int(n_samples), 480000) // 160)
    c = (f - 1) // 2 + 1
    return (c - 4) // 4 + 1


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


class Engine:
    """One model replica on one GPU.  All heavy lifting happens inside the C library."""

    def __init__(self, enc_layers=32, dec_layers=28, mode="bf16", device=0, max_batch=8, max_prompt=448, max_new=256,
                 debug=False):
        self.lib = load_library()
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}")
        self.mode = mode
        self.cfg = SonicConfig(int(device), MODES[mode], int(enc_layers), int(dec_layers), int(max_batch), int(max_prompt),
                               int(max_new), 1 if debug else 0)
        self.h = C.c_void_p()
        if self.lib.sonic_create(C.byref(self.cfg), C.byref(self.h)) != 0:
            msg = self.lib.sonic_last_error(None).decode()
            self.h = None
            raise RuntimeError(msg)
        self.max_batch, self.max_prompt, self.max_new = int(max_batch), int(max_prompt), int(max_new)
        self._last_batch = 0
        self.skipped_tensors = []

    # -- plumbing -----------------------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.sonic_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.sonic_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ------------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: dict):
        """sd: HF-named tensors (torch CPU tensors or numpy arrays), fp32 or bf16 — a dict, or any iterable of
        (name, tensor) pairs (tensors are consumed one at a time, so a generator keeps the host footprint at one tensor)."""
        import torch  # plumbing only: reading checkpoint tensors

        for name, t in (sd.items() if hasattr(sd, "items") else sd):
            if isinstance(t, np.ndarray):
                t = torch.from_numpy(t)
            t = t.detach().cpu().contiguous()
            if t.dtype == torch.bfloat16:
                dt = 1
            else:
                t = t.to(torch.float32)
                dt = 0
            shape = (C.c_int64 * t.dim())(*t.shape)
            rc = self.lib.sonic_load_tensor(self.h, name.encode(), C.c_void_p(t.data_ptr()), dt, shape, t.dim())
            if rc == ERR_UNKNOWN_TENSOR:
                # a real checkpoint may carry tensors the path does not use (rotary inv_freq buffers, tied heads, ...):
                # skip them; sonic_finalize_weights still fails if a tensor the model NEEDS never arrived
                self.skipped_tensors.append(name)
                continue
            self._ck(rc)
        if self.skipped_tensors:
            import warnings
            warnings.warn(f"sonicscribe_b200: ignored {len(self.skipped_tensors)} checkpoint tensors the model does not use: "
                          f"{self.skipped_tensors[:4]}{'...' if len(self.skipped_tensors) > 4 else ''}")
        self._ck(self.lib.sonic_finalize_weights(self.h))

    # -- stages -------------------------------------------------------------------------------------------------------
    @staticmethod
    def _pack(segments: Sequence[np.ndarray], dtype=np.float32):
        """(base address, offsets in samples, lengths, keep-alive list): the C ABI addresses segment b as base[offsets[b] ...],
        so separately allocated host arrays are passed as they are (no concatenation copy)."""
        segs = [np.ascontiguousarray(np.asarray(s, dtype=dtype).reshape(-1)) for s in segments]
        lens = np.array([s.shape[0] for s in segs], dtype=np.int32)
        addrs = [s.ctypes.data for s in segs]
        base = min(addrs)
        isz = np.dtype(dtype).itemsize
        if any((a - base) % isz for a in addrs):               # cannot happen for numpy-allocated arrays; fall back to one copy
            cat = np.concatenate(segs)
            offs = np.zeros(len(segs), dtype=np.int64)
            offs[1:] = np.cumsum(lens[:-1])
            return cat.ctypes.data, offs, lens, [cat]
        offs = np.array([(a - base) // isz for a in addrs], dtype=np.int64)
        return base, offs, lens, segs

    def mel(self, segments, flags=FLAG_REFERENCE_PRESTEP, want_features=True):
        """segments: float32 arrays, or int16 arrays together with FLAG_PCM_S16 (the wire format; widened on the device)."""
        pcm, offs, lens, _keep = self._pack(segments, np.int16 if flags & FLAG_PCM_S16 else np.float32)
        B = len(lens)
        feats = np.empty((B, N_MELS, N_FRAMES), dtype=np.float32) if want_features else None
        nfr = np.zeros(B, dtype=np.int32)
        self._ck(self.lib.sonic_mel(self.h, C.c_void_p(pcm), offs.ctypes.data_as(C.POINTER(C.c_int64)), _i32p(lens), B, flags,
                                    feats.ctypes.data_as(C.c_void_p) if want_features else None, _i32p(nfr)))
        self._last_batch = B
        return feats, nfr

    def encode(self, want_embeds=True):
        B = self._last_batch
        emb = np.empty((B, MERGED, DEC_HIDDEN), dtype=np.float32) if want_embeds else None
        na = np.zeros(B, dtype=np.int32)
        self._ck(self.lib.sonic_encode(self.h, B, _f32p(emb), _i32p(na)))
        return emb, na

    @staticmethod
    def _pack_ids(prompts):
        lens = [len(p) for p in prompts]
        offs = np.zeros(len(prompts) + 1, dtype=np.int32)
        offs[1:] = np.cumsum(lens)
        ids = np.concatenate([np.asarray(p, dtype=np.int32) for p in prompts]).astype(np.int32)
        return ids, offs

    def generate(self, prompts, max_new_tokens, want_margins=False):
        ids, offs = self._pack_ids(prompts)
        B = len(prompts)
        out = np.zeros((B, max_new_tokens), dtype=np.int32)
        n_out = np.zeros(B, dtype=np.int32)
        mar = np.zeros((B, max_new_tokens), dtype=np.float32) if want_margins else None
        self._ck(self.lib.sonic_generate(self.h, _i32p(ids), _i32p(offs), B, max_new_tokens, _i32p(out), _i32p(n_out), _f32p(mar)))
        toks = [out[b, : n_out[b]].tolist() for b in range(B)]
        return (toks, [mar[b, : n_out[b]] for b in range(B)]) if want_margins else toks

    def transcribe_ids(self, segments, prompts, max_new_tokens, flags=FLAG_REFERENCE_PRESTEP, want_margins=False):
        """The whole hot path for a batch of host segments -> generated token ids (one C call)."""
        pcm, offs, lens, _keep = self._pack(segments, np.int16 if flags & FLAG_PCM_S16 else np.float32)
        return self.transcribe_packed(pcm, offs, lens, prompts, max_new_tokens, flags, want_margins)

    def transcribe_packed(self, pcm_ptr, offs, lens, prompts, max_new_tokens, flags=FLAG_REFERENCE_PRESTEP, want_margins=False):
        ids, ioffs = self._pack_ids(prompts)
        B = len(prompts)
        out = np.zeros((B, max_new_tokens), dtype=np.int32)
        n_out = np.zeros(B, dtype=np.int32)
        mar = np.zeros((B, max_new_tokens), dtype=np.float32) if want_margins else None
        self._ck(self.lib.sonic_transcribe_batch(self.h, C.c_void_p(pcm_ptr), offs.ctypes.data_as(C.POINTER(C.c_int64)), _i32p(lens), B, flags,
                                                 _i32p(ids), _i32p(ioffs), max_new_tokens, _i32p(out), _i32p(n_out), _f32p(mar)))
        self._last_batch = B
        toks = [out[b, : n_out[b]].tolist() for b in range(B)]
        return (toks, [mar[b, : n_out[b]] for b in range(B)]) if want_margins else toks

    # -- instrumentation ---------------------------------------------------------------------------------------------
    def sync(self):
        self._ck(self.lib.sonic_sync(self.h))

    def timer_begin(self):
        self._ck(self.lib.sonic_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.sonic_timer_end(self.h, C.byref(ms)))
        return float(ms.value)

    def stage_times(self):
        a = (C.c_float * 4)()
        self._ck(self.lib.sonic_stage_times(self.h, a))
        return dict(zip(("mel_ms", "encode_ms", "prefill_ms", "decode_ms"), [float(v) for v in a]))

    def profile_begin(self):
        self._ck(self.lib.sonic_profile_begin(self.h))

    def profile_end(self) -> dict:
        n = int(self.lib.sonic_profile_num_classes())
        ms = (C.c_float * n)()
        cnt = (C.c_int64 * n)()
        self._ck(self.lib.sonic_profile_end(self.h, ms, cnt, n))
        return {self.lib.sonic_profile_class_name(i).decode(): {"ms": float(ms[i]), "launches": int(cnt[i])} for i in range(n)}

    def launch_count(self) -> int:
        return int(self.lib.sonic_launch_count(self.h))

    def device_bytes(self) -> int:
        return int(self.lib.sonic_device_bytes(self.h))

    def debug_set_logit_steps(self, steps):
        a = np.asarray(list(steps), dtype=np.int32)
        self._ck(self.lib.sonic_debug_set_logit_steps(self.h, _i32p(a) if len(a) else None, len(a)))

    def debug_read(self, name: str, max_elems: int) -> np.ndarray:
        out = np.empty(max_elems, dtype=np.float32)
        n = C.c_size_t()
        self._ck(self.lib.sonic_debug_read(self.h, name.encode(), _f32p(out), max_elems, C.byref(n)))
        return out[: n.value].copy()

    def test_gemm(self, A, W, bias=None, resid=None, act=0, impl=0, swap=False):
        A = np.ascontiguousarray(A, dtype=np.float32)
        W = np.ascontiguousarray(W, dtype=np.float32)
        M, K = A.shape
        N = W.shape[0]
        outN = N // 2 if act == 2 else N
        Cm = np.zeros((M, outN), dtype=np.float32)
        b = np.ascontiguousarray(bias, dtype=np.float32) if bias is not None else None
        r = np.ascontiguousarray(resid, dtype=np.float32) if resid is not None else None
        self._ck(self.lib.sonic_test_gemm(self.h, impl, 1 if swap else 0, _f32p(A), _f32p(W), _f32p(b), _f32p(r), _f32p(Cm), M, N, K, act))
        return Cm

    def test_enc_attention(self, qkv, segments, T, impl=0):
        qkv = np.ascontiguousarray(qkv, dtype=np.float32)
        out = np.zeros((segments * T, 1280), dtype=np.float32)
        self._ck(self.lib.sonic_test_enc_attention(self.h, impl, _f32p(qkv), _f32p(out), segments, T))
        return out

    def bench_gemm(self, M, N, K, swap=False, act=0, iters=50) -> float:
        us = C.c_float()
        self._ck(self.lib.sonic_bench_gemm(self.h, 1 if swap else 0, M, N, K, act, iters, C.byref(us)))
        return float(us.value)

    def bench_mma(self, m, ntok, n_mma=512, n_acc=8, n_tiles=4):
        a, b = C.c_float(), C.c_float()
        self._ck(self.lib.sonic_bench_mma(self.h, m, ntok, n_mma, n_acc, n_tiles, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def test_gemm_int8(self, A, W, bias=None, resid=None, act=0, swap=False):
        A = np.ascontiguousarray(A, dtype=np.float32)
        W = np.ascontiguousarray(W, dtype=np.float32)
        M, K = A.shape
        N = W.shape[0]
        outN = N // 2 if act == 2 else N
        Cm = np.zeros((M, outN), dtype=np.float32)
        b = np.ascontiguousarray(bias, dtype=np.float32) if bias is not None else None
        r = np.ascontiguousarray(resid, dtype=np.float32) if resid is not None else None
        self._ck(self.lib.sonic_test_gemm_int8(self.h, 1 if swap else 0, _f32p(A), _f32p(W), _f32p(b), _f32p(r), _f32p(Cm), M, N, K, act))
        return Cm
