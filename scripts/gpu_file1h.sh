#!/bin/bash
# 1 h file (180 segments), strong scaling over N GPUs of one box
cd "$(dirname "$0")/.."
N=${1:-1}; O=gpurun_out/multi; mkdir -p $O
if [ "$N" = "1" ]; then TR="python bench.py --gpus 1"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"; fi
timeout 600 $TR --workload file1h --steps 2 --warmup 1 > $O/file1h_${N}gpu.json 2> $O/file1h_${N}gpu.err; echo "file1h N=$N rc=$?"; tail -1 $O/file1h_${N}gpu.json | cut -c1-260
