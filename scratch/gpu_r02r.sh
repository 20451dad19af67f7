#!/bin/bash
# r02r: bench line, launch list of one forward, ncu --set full of the bf16 pair+top-k kernel
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r02r.json 2> gpurun_out/bench_r02r.err; tail -c 300 gpurun_out/bench_r02r.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02r.csv python bench.py --profile > gpurun_out/profile_r02r.log 2>&1; tail -2 gpurun_out/profile_r02r.log
cap() { # name regex script args...
  local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/${name}_r02r "$@" > gpurun_out/ncu_${name}_r02r.log 2>&1; tail -1 gpurun_out/ncu_${name}_r02r.log
}
cap pair_topk_bf16_n100 pair_topk python scratch/ppn_prof.py 100 4096 4 bf16
cap pair_topk_bf16_n200 pair_topk python scratch/ppn_prof.py 200 2048 4 bf16
cap pair_topk_bf16_n400 pair_topk python scratch/ppn_prof.py 400 1024 4 bf16
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r02r.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['e2e']['serial']['value'], d['ms_per_step'], d['breakdown_ms'], d['gpu_launches_per_step'])
"
