#!/bin/bash
# round 2 state check: GPU tests, bench, relation-fusion timing, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; tail -c 3000 gpurun_out/bench_r02a.json
timeout 300 python scratch/rel_time.py > gpurun_out/rel_time.log 2>&1; cat gpurun_out/rel_time.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02a.csv python bench.py --profile > gpurun_out/prof.log 2>&1
ls -la gpurun_out
