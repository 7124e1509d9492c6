#!/bin/bash
mkdir -p gpurun_out
echo "== pytest ransac"; timeout 900 python -m pytest tests/test_gpu_ransac.py -m gpu -q --timeout 300 2>&1 | tail -6
for cfg in "2 8 0" "2 8 1" "4 4 1" "2 16 1" "4 8 1"; do set -- $cfg
  echo "== ransac hpt=$1 rounds=$2 packed=$3"; timeout 600 python bench.py --workload ransac --steps 5 --hpt $1 --rounds $2 --packed $3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])
    else: print(l.strip()[:200])"
done
C="python bench.py --workload ransac --pairs 64 --steps 2 --warmup 3 --no-cpu --hpt 2 --packed 1"
timeout 300 $C > gpurun_out/plain_ransac_p.json 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ransac_aca -s 3 -c 1 -f -o gpurun_out/prof_ransac_packed $C > gpurun_out/ncu_ransac_p.log 2>&1; echo "ncu rc=$?"
echo "== zero-copy experiment"; timeout 600 python tools/zerocopy_test.py 2>&1 | tail -8
