#!/bin/bash
# r02aa: final bench line + ncu --set full of the two-group bf16 pair+top-k kernel
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r02aa.json 2> gpurun_out/bench_r02aa.err; tail -c 200 gpurun_out/bench_r02aa.err
cap() { local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/${name}_r02aa "$@" > gpurun_out/ncu_${name}_r02aa.log 2>&1; tail -1 gpurun_out/ncu_${name}_r02aa.log; }
cap pair_topk_bf16_n100 pair_topk python scratch/ppn_prof.py 100 4096 4 bf16
cap pair_topk_bf16_n200 pair_topk python scratch/ppn_prof.py 200 2048 4 bf16
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r02aa.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['breakdown_ms'])
print([ (r['N'], round(r['frac_of_hbm_peak'],3), round(r['bf16']['frac_of_hbm_peak'],3), round(r['bf16']['images_per_sec']/1e6,1)) for r in d['ppn_microbench']])
"
