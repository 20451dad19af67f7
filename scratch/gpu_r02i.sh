#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline --no-ppn-microbench --no-eager-baseline --no-train > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02i.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['breakdown_ms'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['roofline']['frac_of_3xtf32_ceiling'])"; tail -3 gpurun_out/bench_r02i.err
