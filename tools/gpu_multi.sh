#!/bin/bash
# multi-GPU: launched exactly as the driver does
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== bench N=$N default"; timeout 900 $R bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
echo "== reference arm N=$N"; timeout 600 $R bench.py --impl reference --gpus $N --steps 3 --warmup 3 2>/dev/null | tail -2
echo "== ransac N=$N"; timeout 900 $R bench.py --gpus $N --workload ransac --steps 5 > gpurun_out/ransac_n$N.json 2> gpurun_out/ransac_n$N.err; tail -3 gpurun_out/ransac_n$N.err; cat gpurun_out/ransac_n$N.json
echo "== rect N=$N"; timeout 900 $R bench.py --gpus $N --workload rect_f32 --steps 20 > gpurun_out/rect_n$N.json 2> gpurun_out/rect_n$N.err; tail -3 gpurun_out/rect_n$N.err; cat gpurun_out/rect_n$N.json
