#!/bin/bash
mkdir -p gpurun_out
for w in sks_f32 aca_f64 ge_f64 gpt_f64; do
  C="python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu"
  timeout 300 $C > gpurun_out/plain_$w.json 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_aos -s 3 -c 1 -f -o gpurun_out/prof_$w $C > gpurun_out/ncu_$w.log 2>&1
  echo "$w full rc=$?"
done
