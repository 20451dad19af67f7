#!/bin/bash
# final ncu --set full captures of the round-2 kernels
mkdir -p gpurun_out
cap() { # name regex script args...
  local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/${name}_r02l "$@" > gpurun_out/ncu_${name}_r02l.log 2>&1; tail -1 gpurun_out/ncu_${name}_r02l.log
}
cap umma_gemm umma_gemm python scratch/kern_prof.py gemm
cap fa_umma fa_umma python scratch/kern_prof.py fa
cap pair_topk_n100 pair_topk python scratch/ppn_prof.py 100 4096 4
cap pair_topk_n200 pair_topk python scratch/ppn_prof.py 200 2048 4
cap pair_topk_n400 pair_topk python scratch/ppn_prof.py 400 1024 4
cap conv2_umma conv2_umma python scratch/conv_prof.py
