#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "igemm2_pair" > gpurun_out/c40_pair.log 2>&1
echo "pair test rc=$?"; tail -5 gpurun_out/c40_pair.log
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/c40_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/c40_kernels.log
timeout 600 python tests/prof_epi_clocks.py 16 > gpurun_out/c40_clocks.log 2>&1
echo "prof rc=$?"; grep pairs gpurun_out/c40_clocks.log | cut -c1-110
