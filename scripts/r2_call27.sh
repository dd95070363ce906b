#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/c27_kernels.log 2>&1
echo "kernels rc=$?"; tail -4 gpurun_out/c27_kernels.log
timeout 600 python tests/prof_rings.py 16 > gpurun_out/c27_rings.log 2>&1
echo "prof rc=$?"; cat gpurun_out/c27_rings.log
timeout 600 python tests/prof_kernels.py > gpurun_out/c27_prof_kernels.log 2>&1
echo "prof_kernels rc=$?"; grep -i "conv1\|pack\|unpack" gpurun_out/c27_prof_kernels.log | head
