#!/bin/bash
mkdir -p gpurun_out
free -g | head -2
echo "== pytest ransac"; timeout 900 python -m pytest tests/test_gpu_ransac.py -m gpu -q --timeout 300 2>&1 | tail -3
for cfg in "2 8 1" "4 4 1" "2 8 0" "4 4 5"; do set -- $cfg
  echo "== ransac hpt=$1 rounds=$2 packed/threads=$3"; timeout 600 python bench.py --workload ransac --steps 5 --hpt $1 --rounds $2 --packed $3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])
    else: print(l.strip()[:200])"
done
