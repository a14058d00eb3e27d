#!/bin/bash
# multi-GPU lines: 1 h file (180 segments) strong scaling and the default weak-scaling bench at N GPUs
cd "$(dirname "$0")/.."
N=${1:-2}; O=gpurun_out/multi; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"
timeout 600 $TR --workload file1h --steps 2 --warmup 1 > $O/file1h_${N}gpu.json 2> $O/file1h_${N}gpu.err; echo "file1h rc=$?"; tail -1 $O/file1h_${N}gpu.json | cut -c1-700
timeout 600 $TR --steps 5 --warmup 3 --no-cpu-baseline --no-api-threads > $O/default_${N}gpu.json 2> $O/default_${N}gpu.err; echo "default rc=$?"; tail -1 $O/default_${N}gpu.json | cut -c1-400
