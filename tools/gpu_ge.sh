#!/bin/bash
mkdir -p gpurun_out
echo "== GE + reference-kernel pins"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "ge_competitor or without_fma" 2>&1 | tail -5
for w in ge_f32 ge_f64; do echo "== bench $w"; timeout 600 python bench.py --workload $w 2>gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json | cut -c1-1500; done
