#!/bin/bash
# compute-sanitizer memcheck + racecheck on the shared-memory / TMA / RANSAC kernels at small sizes
# (SURVEY.md section 5).  Logs (or the tool's refusal, verbatim) go to gpurun_out/sanitizer_*.log.
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 400 $CS --tool $tool --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.log
  tail -12 gpurun_out/sanitizer_$tool.log
done
