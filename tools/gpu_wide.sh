#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8
echo "== hint sweep"; timeout 900 python tools/hint_sweep.py 2>&1 | tee gpurun_out/hint_sweep.log | tail -40
echo "== sweep quick"; timeout 900 python tools/sweep.py --quick 2>&1 | grep -E "direct|soa" | tee gpurun_out/sweep_wide.log
