#!/bin/bash
# round 1, second session: 8-GPU numbers with the current RANSAC scorer and the NUMA-bound host path
N=${1:-8}
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
(nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; nproc) > gpurun_out/topo_n$N.txt 2>&1
echo "== bench default N=$N (NUMA bound)"; timeout 600 $R bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -2 gpurun_out/bench_n$N.err; cut -c1-2200 gpurun_out/bench_n$N.json
echo "== bench default N=$N (not bound)"; timeout 600 $R bench.py --gpus $N --steps 5 --warmup 3 --no-numa-bind > gpurun_out/bench_n${N}_nonuma.json 2> gpurun_out/bench_n${N}_nonuma.err; tail -2 gpurun_out/bench_n${N}_nonuma.err; cut -c1-2200 gpurun_out/bench_n${N}_nonuma.json
echo "== ransac NCCL N=$N"; timeout 600 $R bench.py --gpus $N --workload ransac --steps 10 > gpurun_out/ransac_n$N.json 2> gpurun_out/ransac_n$N.err; tail -2 gpurun_out/ransac_n$N.err; cut -c1-1800 gpurun_out/ransac_n$N.json
echo "== ransac peer N=$N"; timeout 600 $R bench.py --gpus $N --workload ransac --steps 10 --peer-reduce --no-cpu > gpurun_out/ransac_peer_n$N.json 2> gpurun_out/ransac_peer_n$N.err; tail -2 gpurun_out/ransac_peer_n$N.err; cut -c1-600 gpurun_out/ransac_peer_n$N.json
echo "== reference arm N=$N"; timeout 600 $R bench.py --impl reference --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_ref_n$N.json 2>/dev/null; cut -c1-500 gpurun_out/bench_ref_n$N.json
