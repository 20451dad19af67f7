#!/bin/bash
# bench + reference arm + launch list + full captures for profiles/
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -c 600 gpurun_out/bench_b.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_b.json 2>> gpurun_out/bench_b.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01g.csv python bench.py --profile > gpurun_out/prof.log 2>&1
for k in gemm:umma_gemm fa:fa_umma pair:pair_umma; do
  w=${k%%:*}; r=${k##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$r -s 2 -c 1 -f -o gpurun_out/${r}_r01g python scratch/kern_prof.py $w > gpurun_out/ncu_$w.log 2>&1
done
ls -la gpurun_out | head -40
