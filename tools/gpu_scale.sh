#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
echo "== bench default N=$N"; timeout 600 $R bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; cut -c1-330 gpurun_out/bench_n$N.json
echo "== ransac N=$N"; timeout 600 $R bench.py --gpus $N --workload ransac --steps 10 --no-cpu > gpurun_out/ransac_n$N.json 2> gpurun_out/ransac_n$N.err; cut -c1-330 gpurun_out/ransac_n$N.json
echo "== gloo-free dist tests on GPUs: 2-rank sharded solve"; timeout 300 $R tools/peer_reduce_test.py 2>&1 | tail -2
