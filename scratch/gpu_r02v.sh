#!/bin/bash
# r02v: final state -- GPU tests, bench line, reference arm, launch list of one forward
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r02v.json 2> gpurun_out/bench_r02v.err; tail -c 300 gpurun_out/bench_r02v.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02v.json 2> gpurun_out/bench_ref_r02v.err; tail -c 400 gpurun_out/bench_ref_r02v.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02v.csv python bench.py --profile > gpurun_out/profile_r02v.log 2>&1; tail -2 gpurun_out/profile_r02v.log
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r02v.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['e2e']['serial']['value'], d['ms_per_step'], d['breakdown_ms'], d['gpu_launches_per_step'])
print(d['gpu_eager_baseline'])
print(d['config4_e2e']); print(d['config4_head']); print(d['e2e_simple_test']['value']); print(d['reduced_precision_arm']['value']); print(d['cpu_baseline'])
"
