#!/bin/bash
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
echo "== ransac tests"; timeout 600 python -m pytest tests/test_gpu_ransac.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -k "ransac or pair" 2>&1 | tail -2
echo "== pairs N=2"; timeout 600 $R bench.py --gpus 2 --workload ransac --ransac-shard pairs --steps 10 --no-cpu > gpurun_out/ransac_pairs_n2.json 2> gpurun_out/ransac_pairs_n2.err; tail -1 gpurun_out/ransac_pairs_n2.err; cut -c1-700 gpurun_out/ransac_pairs_n2.json
