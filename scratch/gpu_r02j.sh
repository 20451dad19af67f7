#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 2 -c 1 -f -o gpurun_out/umma_gemm_r02j python scratch/kern_prof.py gemm > gpurun_out/ncu_gemm_r02j.log 2>&1; tail -2 gpurun_out/ncu_gemm_r02j.log
