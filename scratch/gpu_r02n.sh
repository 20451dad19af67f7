#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s -k "single_pass or full_size or chain" > gpurun_out/pytest_sp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sp.log
grep -n "single-pass\|passed\|failed\|rc=" gpurun_out/pytest_sp.log | tail -8
timeout 900 python bench.py --no-cpu-baseline --no-ppn-microbench --no-eager-baseline --no-train > gpurun_out/bench_r02n.json 2> gpurun_out/bench_r02n.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02n.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d.get('reduced_precision_arm'))"; tail -3 gpurun_out/bench_r02n.err
