#!/bin/bash
# r02y: ncu --set full of the FINAL 3xBF16 GEMM (split rings, 8 epilogue warps) on the FFN1 / FFN2 problems, and of the skinny FFN1 GEMM
mkdir -p gpurun_out
cap() { # name regex script args...
  local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/${name}_r02y "$@" > gpurun_out/ncu_${name}_r02y.log 2>&1; tail -1 gpurun_out/ncu_${name}_r02y.log
}
cap umma_gemm_bf16x3_ffn1 umma_gemm python scratch/kern_prof.py gemm16
cap umma_gemm_bf16x3_ffn2 umma_gemm python scratch/kern_prof.py gemm16_ffn2
