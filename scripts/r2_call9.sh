#!/bin/bash
set -x
mkdir -p gpurun_out
python tests/prof_step.py 3 16 > gpurun_out/c9_plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_batch16_v2.csv python tests/prof_step.py 3 16 > gpurun_out/c9_ncu_step.log 2>&1
python tests/prof_kernels.py wg42 0 3 > gpurun_out/c9_plain_wg42.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_wgrad256_persistent_conv42c python tests/prof_kernels.py wg42 0 3 > gpurun_out/c9_ncu1.log 2>&1
python tests/prof_kernels.py wg22 4 3 > gpurun_out/c9_plain_wg22.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wgrad5_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_wgrad5_conv22c python tests/prof_kernels.py wg22 4 3 > gpurun_out/c9_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:unpack_all_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_unpack_all python tests/prof_step.py 3 16 > gpurun_out/c9_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv1_wgrad_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_conv1_wgrad python tests/prof_step.py 3 16 > gpurun_out/c9_ncu4.log 2>&1
cat gpurun_out/c9_plain_*.log
timeout 300 python -m pytest tests/test_gpu_dropin.py -x -q > gpurun_out/c9_pytest_dropin.log 2>&1; tail -3 gpurun_out/c9_pytest_dropin.log
