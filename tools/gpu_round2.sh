#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
echo "== bench default"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
echo "== bench ransac small"; timeout 600 python bench.py --workload ransac --pairs 64 --steps 3 > gpurun_out/ransac_small.json 2> gpurun_out/ransac_small.err; tail -3 gpurun_out/ransac_small.err; cat gpurun_out/ransac_small.json
echo "== bench ransac full"; timeout 900 python bench.py --workload ransac --steps 5 > gpurun_out/ransac_n1.json 2> gpurun_out/ransac_n1.err; tail -3 gpurun_out/ransac_n1.err; cat gpurun_out/ransac_n1.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 | tee gpurun_out/bench_ref.json
for w in sks_f32 rect_f32 aca_f64 sks_f64; do timeout 600 python bench.py --workload $w --steps 10 --no-gpu-baseline | tee gpurun_out/bench_$w.json; done
nproc; lscpu | head -20
