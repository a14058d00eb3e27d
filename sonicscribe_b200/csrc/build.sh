#!/bin/bash
# Build libsonic_b200.so in-tree for sm_100a.  Usage: build.sh [outdir]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${1:-$HERE/..}"
OBJ="$HERE/build"
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
pids=()
for f in mel gemm_simt ops attention gemm_tc attention_tc attention_prefill_tc decode_attn decode_persist decode_rs api; do
  if [ ! -f "$OBJ/$f.o" ] || [ "$HERE/$f.cu" -nt "$OBJ/$f.o" ] || [ -n "$(find "$HERE" -maxdepth 1 \( -name '*.h' -o -name '*.cuh' -o -name '*.inc' \) -newer "$OBJ/$f.o" 2>/dev/null)" ] || [ "$HERE/../../include/sonic_b200.h" -nt "$OBJ/$f.o" ]; then
    $NVCC $FLAGS -c "$HERE/$f.cu" -o "$OBJ/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o "$OUT/libsonic_b200.so" "$OBJ"/*.o -cudart static -Xlinker --no-undefined -lpthread -ldl -lrt
echo "built $OUT/libsonic_b200.so"
