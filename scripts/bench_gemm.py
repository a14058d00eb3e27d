import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sonicscribe_b200.engine import Engine
eng = Engine(1, 1, mode="bf16", device=0, max_batch=1, max_prompt=64, max_new=8)
shapes = {"qkv": (3072, 2048, 0), "o": (2048, 2048, 0), "gateup": (12288, 2048, 2), "down": (2048, 6144, 0), "lmhead": (59264, 2048, 0)}
for M in (16, 64):
    for name, (N, K, act) in shapes.items():
        res = []
        for sp in ([0, 1, 2, 4, 8, 16] if name != "lmhead" else [0]):
            os.environ["SONIC_SPLITS"] = str(sp)   # 0 = library heuristic
            us = eng.bench_gemm(M, N, K, swap=True, act=act, iters=40)
            res.append(f"s{sp}:{us:6.1f}us({N*K*2/us/1e3:5.0f}GB/s)")
        print(f"M={M:3d} {name:7s}", " ".join(res), flush=True)
os.environ.pop("SONIC_SPLITS")
for (M, N, K) in [(24000, 1280, 1280), (24000, 3840, 1280), (24000, 5120, 1280), (24000, 1280, 5120), (4320, 2048, 2048), (4320, 12288, 2048)]:
    us = eng.bench_gemm(M, N, K, swap=False, iters=10)
    print(f"normal M={M} N={N} K={K}: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)
