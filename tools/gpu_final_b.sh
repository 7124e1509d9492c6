#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== reference arm"; timeout 600 python bench.py --impl reference 2>/dev/null | tee gpurun_out/bench_ref_final.json | cut -c1-400
echo "== default bench"; timeout 900 python bench.py 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | cut -c1-1200
