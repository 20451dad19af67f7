#!/bin/bash
# r02w: compute-sanitizer over the kernels added in the last session
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scratch/sanitize_small.py r02v > gpurun_out/sanitizer_r02w_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_r02w_$tool.log
done
