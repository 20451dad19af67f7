#!/bin/bash
# r02t: ncu --set full of the 3xBF16 GEMM (FFN1 / FFN2 problems) + launch list of one forward
mkdir -p gpurun_out
cap() { # name regex script args...
  local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/${name}_r02t "$@" > gpurun_out/ncu_${name}_r02t.log 2>&1; tail -1 gpurun_out/ncu_${name}_r02t.log
}
cap umma_gemm_bf16x3_ffn1 umma_gemm python scratch/kern_prof.py gemm16
cap umma_gemm_bf16x3_ffn2 umma_gemm python scratch/kern_prof.py gemm16_ffn2
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02t.csv python bench.py --profile > gpurun_out/profile_r02t.log 2>&1; tail -2 gpurun_out/profile_r02t.log
