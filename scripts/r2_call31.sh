#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "pack or unpack or sgd" > gpurun_out/c31_kernels.log 2>&1
echo "kernels rc=$?"; tail -4 gpurun_out/c31_kernels.log
timeout 600 python tests/prof_hbm.py 16 > gpurun_out/c31_hbm_new.log 2>&1
echo "prof rc=$?"; cat gpurun_out/c31_hbm_new.log | tail -8
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -m gpu > gpurun_out/c31_net.log 2>&1
echo "net rc=$?"; tail -3 gpurun_out/c31_net.log
