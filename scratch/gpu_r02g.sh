#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_train.log
tail -30 gpurun_out/pytest_train.log
timeout 900 python bench.py --no-cpu-baseline --no-ppn-microbench --no-eager-baseline > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02g.json').read().strip().splitlines()[-1]); print(d['value'], json.dumps(d.get('train_step'), indent=1))"; tail -5 gpurun_out/bench_r02g.err
