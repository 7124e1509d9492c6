#!/bin/bash
# round 1, second session: full GPU suite, RANSAC bench (1 GPU), default bench
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -5
echo "== ransac bench"; timeout 600 python bench.py --workload ransac --steps 5 2>gpurun_out/ransac_v3.err | tee gpurun_out/ransac_v3.json
echo "== default bench"; timeout 900 python bench.py 2>gpurun_out/bench_v3.err | tee gpurun_out/bench_v3.json
