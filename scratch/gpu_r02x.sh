#!/bin/bash
# r02x: final verification -- GPU tests, smoke, bench line, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r02x.json 2> gpurun_out/bench_r02x.err; tail -c 300 gpurun_out/bench_r02x.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02x.csv python bench.py --profile > gpurun_out/profile_r02x.log 2>&1; tail -2 gpurun_out/profile_r02x.log
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r02x.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['e2e']['serial']['value'], d['ms_per_step'], d['breakdown_ms'], d['gpu_launches_per_step'])
print(d['gpu_eager_baseline']['head_b200_ms'], d['gpu_eager_baseline']['head_speedup_vs_eager_fp32'], d['e2e_simple_test']['value'], d['config4_e2e']['graph_replay'], d['clocks'])
print([ (r['N'], round(r['frac_of_hbm_peak'],3), round(r['bf16']['frac_of_hbm_peak'],3)) for r in d['ppn_microbench']])
print({k:(v['ms_per_step']) for k,v in d['train_step'].items() if isinstance(v,dict)})
r=d['roofline']; print(r['achieved'], r['frac'], r['frac_of_3xbf16_ceiling'], r['l2_to_smem_tbs'], r['ms_per_launch'])
"
