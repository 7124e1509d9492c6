#!/bin/bash
# Round-end evidence run on one GPU: tests, smoke, bench lines, ncu launch list + full captures.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
B="python bench.py --steps 20 --warmup 3"
echo "== bench"; timeout 900 $B > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err && cat gpurun_out/bench_n1.json &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 | tee gpurun_out/bench_ref.json
echo "== ransac"; timeout 900 python bench.py --workload ransac --steps 5 | tee gpurun_out/ransac_n1.json
for w in sks_f32 rect_f32 aca_f64 sks_f64; do timeout 600 python bench.py --workload $w --steps 20 --no-gpu-baseline > gpurun_out/bench_$w.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'])"; done
for cfg in "aca_f32 1" "sks_f64 1" "rect_f32 1"; do
  set -- $cfg
  C="python bench.py --workload $1 --variant $2 --steps 3 --warmup 3 --no-e2e --no-cpu"
  timeout 300 $C > gpurun_out/plain_$1_v$2.json 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_aos -s 3 -c 2 -f -o gpurun_out/prof_$1_v$2 $C > gpurun_out/ncu_$1_v$2.log 2>&1
  echo "$cfg full rc=$?"
done
C="python bench.py --workload ransac --pairs 64 --steps 2 --warmup 3 --no-cpu"
timeout 300 $C > gpurun_out/plain_ransac.json 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ransac_aca -s 3 -c 1 -f -o gpurun_out/prof_ransac $C > gpurun_out/ncu_ransac.log 2>&1; echo "ransac ncu rc=$?"
