#!/bin/bash
# ncu evidence: launch list of the bench command + full-set captures of the top kernels.
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3"
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
echo "== plain bench"; timeout 600 $B > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && cat gpurun_out/bench_plain.json &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for cfg in "aca_f32 1" "aca_f32 2" "sks_f64 1" "rect_f32 1"; do
  set -- $cfg
  C="python bench.py --workload $1 --variant $2 --steps 3 --warmup 3 --no-e2e --no-cpu"
  timeout 300 $C > gpurun_out/plain_$1_v$2.json 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_aos -s 3 -c 2 -f -o gpurun_out/prof_$1_v$2 $C > gpurun_out/ncu_$1_v$2.log 2>&1
  echo "$cfg full rc=$?"
done
ls -la gpurun_out
