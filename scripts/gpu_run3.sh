#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c; mkdir -p $O
timeout 200 python scripts/rs_phases.py bf16 1 > $O/rs_dbg_cta0.txt 2>&1; tail -9 $O/rs_dbg_cta0.txt
SONIC_RS_DBG_CTA=147 timeout 200 python scripts/rs_phases.py bf16 1 16 > $O/rs_dbg_cta147.txt 2>&1; tail -18 $O/rs_dbg_cta147.txt
