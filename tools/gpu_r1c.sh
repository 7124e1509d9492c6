#!/bin/bash
mkdir -p gpurun_out
echo "== warp tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "warp_grid" 2>&1 | tail -4
echo "== warp bench"; timeout 300 python tools/warp_bench.py 2>&1 | tee gpurun_out/warp_bench.log
echo "== launch list of the default bench command"
B="python bench.py --steps 5 --warmup 3"
timeout 600 $B > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
echo "== ransac full capture"
C="python bench.py --workload ransac --pairs 64 --steps 3 --warmup 3 --no-cpu --no-e2e"
timeout 300 $C > gpurun_out/plain_ransac_v3.json 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ransac_aca -s 3 -c 1 -f -o gpurun_out/prof_ransac_v3 $C > gpurun_out/ncu_ransac_v3.log 2>&1
echo "ransac full rc=$?"
echo "== ge full capture"
C="python bench.py --workload ge_f32 --steps 3 --warmup 3 --no-e2e --no-cpu"
timeout 300 $C > gpurun_out/plain_ge_f32.json 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_aos -s 3 -c 1 -f -o gpurun_out/prof_ge_f32 $C > gpurun_out/ncu_ge_f32.log 2>&1
echo "ge full rc=$?"
ls -la gpurun_out | tail -8
