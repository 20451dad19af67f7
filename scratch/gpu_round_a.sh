#!/bin/bash
# tests + bench + launch list + full captures of the three tcgen05 kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 3000 gpurun_out/bench_a.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01f.csv python bench.py --profile > gpurun_out/prof.log 2>&1
for k in gemm:umma_gemm fa:fa_umma pair:pair_umma; do
  w=${k%%:*}; r=${k##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$r -s 2 -c 1 -f -o gpurun_out/${r}_r01f python scratch/kern_prof.py $w > gpurun_out/ncu_$w.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:topk -s 2 -c 2 -f -o gpurun_out/topk_r01f python scratch/kern_prof.py pair > gpurun_out/ncu_topk.log 2>&1
ls -la gpurun_out
