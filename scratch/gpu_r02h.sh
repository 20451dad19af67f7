#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 2 -c 1 -f -o gpurun_out/umma_gemm_r02h python scratch/kern_prof.py gemm > gpurun_out/ncu_gemm_r02h.log 2>&1; tail -2 gpurun_out/ncu_gemm_r02h.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_umma -s 2 -c 1 -f -o gpurun_out/fa_umma_r02h python scratch/kern_prof.py fa > gpurun_out/ncu_fa_r02h.log 2>&1; tail -2 gpurun_out/ncu_fa_r02h.log
