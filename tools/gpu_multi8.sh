#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== peer reduce test N=$N"; timeout 300 $R tools/peer_reduce_test.py 2>&1 | tail -2
echo "== ransac NCCL N=$N"; timeout 600 $R bench.py --gpus $N --workload ransac --steps 10 > gpurun_out/ransac_n$N.json 2> gpurun_out/ransac_n$N.err; tail -2 gpurun_out/ransac_n$N.err; cat gpurun_out/ransac_n$N.json
echo "== ransac peer N=$N"; timeout 600 $R bench.py --gpus $N --workload ransac --steps 10 --peer-reduce > gpurun_out/ransac_peer_n$N.json 2> gpurun_out/ransac_peer_n$N.err; tail -2 gpurun_out/ransac_peer_n$N.err; cat gpurun_out/ransac_peer_n$N.json
echo "== rect strong 2^28 N=$N"; timeout 600 $R bench.py --gpus $N --workload rect_f32 --log2n 28 --strong --steps 20 > gpurun_out/rect_strong_n$N.json 2> gpurun_out/rect_strong_n$N.err; tail -2 gpurun_out/rect_strong_n$N.err; cat gpurun_out/rect_strong_n$N.json
echo "== bench default N=$N"; timeout 600 $R bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -2 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
