#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -q -s -k "conv_tiny" > gpurun_out/pytest_conv.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_conv.log
tail -30 gpurun_out/pytest_conv.log
timeout 300 python scratch/conv_time.py > gpurun_out/conv_time.log 2>&1; cat gpurun_out/conv_time.log
