#!/bin/bash
# full GPU suite + bench (new legs) + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err; tail -c 6000 gpurun_out/bench_r02e.json; tail -5 gpurun_out/bench_r02e.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02e.csv python bench.py --profile > gpurun_out/prof.log 2>&1
