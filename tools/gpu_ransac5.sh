#!/bin/bash
mkdir -p gpurun_out
echo "== pytest ransac"; timeout 900 python -m pytest tests/test_gpu_ransac.py -m gpu -q --timeout 300 2>&1 | tail -3
echo "== sweep"; timeout 600 python tools/ransac_sweep.py 2>&1 | tee gpurun_out/ransac_sweep4.log
echo "== bench"; timeout 600 python bench.py --workload ransac --steps 5 2>gpurun_out/ransac_v2.err | tee gpurun_out/ransac_v2.json
