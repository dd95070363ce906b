#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "dgrad or upconv or pair or conv3x3" > gpurun_out/c25_kernels.log 2>&1
echo "kernels rc=$?"; tail -4 gpurun_out/c25_kernels.log
timeout 600 python tests/prof_streamk.py 16 > gpurun_out/c25_streamk.log 2>&1
echo "prof rc=$?"; cat gpurun_out/c25_streamk.log
