#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 2 -c 1 -f -o gpurun_out/umma_gemm_r02a python scratch/kern_prof.py gemm > gpurun_out/ncu_gemm_r02a.log 2>&1
tail -3 gpurun_out/ncu_gemm_r02a.log
ls -la gpurun_out/*.ncu-rep | tail -3
