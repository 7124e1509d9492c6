#!/bin/bash
mkdir -p gpurun_out
echo "== full-size tests"; timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 600 2>&1 | tail -15
echo "== memcheck (small cases)"; timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ransac.py -m gpu -q --timeout 900 -k "ragged or golden or rect or variants or seeded" 2>&1 | tail -12
