#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "dgrad or upconv or pair" > gpurun_out/c24_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/c24_kernels.log
timeout 600 python tests/prof_pair256.py 16 > gpurun_out/c24_pair256.log 2>&1
echo "prof rc=$?"; cat gpurun_out/c24_pair256.log
