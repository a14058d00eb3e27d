import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sonicscribe_b200.engine import Engine
eng = Engine(1, 1, mode="bf16", device=0, max_batch=1, max_prompt=64, max_new=8)
name = sys.argv[1]
shapes = {"qkv": (3072, 2048, 0), "o": (2048, 2048, 0), "gateup": (12288, 2048, 2), "down": (2048, 6144, 0), "lmhead": (59264, 2048, 0)}
N, K, act = shapes[name]
M = int(sys.argv[2])
print(eng.bench_gemm(M, N, K, swap=True, act=act, iters=int(sys.argv[3])))
