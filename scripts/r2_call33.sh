#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "pack or unpack or sgd" > gpurun_out/c33_kernels.log 2>&1
echo "kernels rc=$?"; tail -4 gpurun_out/c33_kernels.log
timeout 600 python tests/prof_hbm.py 16 > gpurun_out/c33_hbm.log 2>&1
echo "prof rc=$?"; head -5 gpurun_out/c33_hbm.log; tail -1 gpurun_out/c33_hbm.log
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -m gpu -k "reproducible or parity" > gpurun_out/c33_net.log 2>&1
echo "net rc=$?"; tail -3 gpurun_out/c33_net.log
