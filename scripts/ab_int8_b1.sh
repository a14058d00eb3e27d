#!/bin/bash
# A/B of the batch-1 decode step across library builds (SONIC_LIB).  Older builds are made with
#   mkdir -p .scratch/<commit> && git archive <commit> sonicscribe_b200/csrc include | tar -x -C .scratch/<commit> && bash .scratch/<commit>/sonicscribe_b200/csrc/build.sh
# (this is how the 18 % int8 regression of commit f0f0b9a was traced to the shared-memory carve-out)
for c in b1973f4 2c7727a 86dd6f9 HEAD; do
  if [ "$c" = HEAD ]; then unset SONIC_LIB; else export SONIC_LIB=$PWD/.scratch/$c/sonicscribe_b200/libsonic_b200.so; fi
  for m in int8 bf16; do
  timeout 200 python bench.py --batch 1 --mode $m --no-cpu-baseline --no-api-threads --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c $m', round(d['decode']['ms_per_token_step'],4), d['stage_ms_last_step'])"
  done
done
