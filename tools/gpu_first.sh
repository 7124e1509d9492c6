#!/bin/bash
# First GPU pass: smoke, parity tests, sweep, bench.  Every step under its own timeout.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 900 python tools/sweep.py 2>&1 | tee gpurun_out/sweep.log | tail -80
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
