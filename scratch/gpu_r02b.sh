#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -x -q -k "pair_matrix or pair_topk or ppn_l2 or topk" > gpurun_out/pytest_ppn.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ppn.log
tail -5 gpurun_out/pytest_ppn.log
timeout 300 python scratch/ppn_fused_time.py > gpurun_out/ppn_fused_time.log 2>&1; cat gpurun_out/ppn_fused_time.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_topk -s 1 -c 1 -f -o gpurun_out/pair_topk_r02b python scratch/ppn_prof.py 100 4096 > gpurun_out/ncu_pair_topk.log 2>&1; tail -3 gpurun_out/ncu_pair_topk.log
