#!/bin/bash
mkdir -p gpurun_out
for cfg in "200 2048" "400 1024"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_topk -s 1 -c 1 -f -o gpurun_out/pair_topk_n$1 python scratch/ppn_prof.py $1 $2 > gpurun_out/ncu_pair_topk_n$1.log 2>&1; tail -2 gpurun_out/ncu_pair_topk_n$1.log
done
