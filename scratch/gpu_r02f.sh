#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; tail -c 7000 gpurun_out/bench_r02f.json; tail -5 gpurun_out/bench_r02f.err
