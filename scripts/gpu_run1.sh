#!/bin/bash
# round-2 GPU run 1: new host-side tests + the extra bench workloads
mkdir -p gpurun_out/r2a
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a/bench_default.json 2> gpurun_out/r2a/bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --workload realtime --steps 3 --warmup 3 > gpurun_out/r2a/bench_realtime.json 2> gpurun_out/r2a/bench_realtime.err; echo "realtime rc=$?"
timeout 400 python bench.py --workload file1h --steps 2 --warmup 1 > gpurun_out/r2a/bench_file1h_1gpu.json 2> gpurun_out/r2a/bench_file1h_1gpu.err; echo "file1h rc=$?"
timeout 400 python bench.py --mode int8 --batch 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/bench_int8_b1.json 2> gpurun_out/r2a/bench_int8_b1.err; echo "int8 rc=$?"
timeout 400 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/bench_bf16_b1.json 2> gpurun_out/r2a/bench_bf16_b1.err; echo "b1 rc=$?"
tail -c 600 gpurun_out/r2a/bench_default.json
