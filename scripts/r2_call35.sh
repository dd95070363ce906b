#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv3x3 or guard or strided" > gpurun_out/c35_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/c35_kernels.log
timeout 600 python tests/prof_shallow.py 16 > gpurun_out/c35_shallow.log 2>&1
echo "prof rc=$?"; cat gpurun_out/c35_shallow.log
