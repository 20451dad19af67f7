#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02k.csv python bench.py --profile > gpurun_out/prof.log 2>&1; tail -2 gpurun_out/prof.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scratch/sanitize_small.py all > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/sanitizer_$tool.log; tail -6 gpurun_out/sanitizer_$tool.log
done
