"""tcgen05.mma rate table: clocks per MMA (issue loop / until complete) for the operand shapes the decode kernels use."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sonicscribe_b200.engine import Engine
eng = Engine(2, 2, mode="fp32", device=0, max_batch=1, max_prompt=16, max_new=4)
print("   M  N(tok) acc tiles   issue clk/MMA  complete clk/MMA")
for m in (64, 128):
    for ntok in (16, 32, 64, 128, 256):
        for n_acc in (1, 2, 8):
            if n_acc * ntok > 512:
                continue
            for n_tiles in (1, 4):
                a, b = eng.bench_mma(m, ntok, 1024, n_acc, n_tiles)
                print(f"{m:4d} {ntok:6d} {n_acc:4d} {n_tiles:5d}   {a:10.1f}   {b:10.1f}", flush=True)
eng.close()
