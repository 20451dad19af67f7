#!/bin/bash
# final validation: tests + smoke + bench + launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; tail -3 gpurun_out/pytest_gpu_c.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_c.log 2>&1; tail -2 gpurun_out/smoke_c.log
timeout 600 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_c.err; tail -c 400 gpurun_out/bench_b.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01h.csv python bench.py --profile > gpurun_out/prof.log 2>&1
ls gpurun_out | head -50
