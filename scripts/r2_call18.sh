#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "igemm_pair" > gpurun_out/c18_pytest.log 2>&1; tail -12 gpurun_out/c18_pytest.log | cut -c1-250
timeout 300 python tests/prof_bn.py 16 2>&1 | tail -12
