#!/bin/bash
# One lease script for every GPU session of this repo: gpurun -- 'bash tools/gpu_session.sh <stage> ...'
# Every stage writes its logs under gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
for stage in "$@"; do
case "$stage" in
  ransac_tests) timeout 900 python -m pytest tests/test_gpu_ransac.py -m gpu -q -x --timeout 600 2>&1 | tail -5 | tee gpurun_out/ransac_tests.log ;;
  ransac_sweep) timeout 900 python tools/ransac_sweep.py 2>&1 | tee gpurun_out/ransac_sweep.log ;;
  tests)        timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -8 | tee gpurun_out/tests.log ;;
  smoke)        timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log ;;
  bench)        timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json ;;
  bench_ref)    timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json ;;
  sanitizer)    timeout 900 bash tools/sanitize.sh 2>&1 | tail -40 ;;
  torch_ref)    timeout 600 python tools/torch_reference_gpu.py 2>&1 | tee gpurun_out/torch_reference_gpu.log ;;
  bench_ransac) timeout 900 python bench.py --workload ransac --steps 5 > gpurun_out/bench_ransac.json 2> gpurun_out/bench_ransac.err; tail -c 2500 gpurun_out/bench_ransac.json ;;
  ncu_ransac)   for cfg in "3 2" "3 4" "1 2"; do set -- $cfg; timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ransac_aca -s 1 -c 1 -f -o gpurun_out/ransac_m$1_h$2 python tools/ransac_once.py 444 $1 $2 > gpurun_out/ncu_ransac_m$1_h$2.log 2>&1; tail -2 gpurun_out/ncu_ransac_m$1_h$2.log; done ;;
  multi_tests)  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 300 2>&1 | tail -15 | tee gpurun_out/multi_tests.log ;;
  cpp_multi)    g++ -std=c++17 -O2 -I include tests/cpp/ransac_multi_main.cpp -o /tmp/ransac_multi -L sks_homography_b200 -lsks_cuda -Wl,-rpath,$PWD/sks_homography_b200 && timeout 300 /tmp/ransac_multi 1024 4096 65536 2>&1 | tee gpurun_out/cpp_multi.log ;;
  bench_n)      N=${NGPU:-2}; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1500 gpurun_out/bench_n$N.json; grep "\[bench\]" gpurun_out/bench_n$N.err ;;
  launch_list)  timeout 600 python bench.py --steps 20 --warmup 5 --sustained-s 0 > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 5 --sustained-s 0 > gpurun_out/ncu_launch.log 2>&1; tail -3 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches_bench.csv ;;
  pcie_multi)   N=${NGPU:-2}; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 tools/pcie_multi_probe.py 2> gpurun_out/pcie_multi_n$N.err | tee gpurun_out/pcie_multi_probe_n$N.log; tail -3 gpurun_out/pcie_multi_n$N.err ;;
  wring_sweep)  timeout 900 python tools/wring_sweep.py 2>&1 | tee gpurun_out/wring_sweep.log ;;
  parity_tests) timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -5 | tee gpurun_out/parity_tests.log ;;
  ncu_headline) timeout 300 python tools/headline_once.py > gpurun_out/headline_plain.log 2>&1 && timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_aos_direct -s 2 -c 1 -f -o gpurun_out/headline python tools/headline_once.py > gpurun_out/ncu_headline.log 2>&1; tail -2 gpurun_out/ncu_headline.log ;;
  *) echo "unknown stage $stage" ;;
esac
done
