#!/bin/bash
# up-convolution 2 alone under ncu --set full: one forward, one dgrad, one wgrad launch (what bounds them?)
mkdir -p gpurun_out
timeout 300 python tests/prof_upconv.py 2 16 > gpurun_out/c50_up2_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/c50_up2_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"igemm_pair_kernel|wgrad" -s 9 -c 15 -f -o gpurun_out/r2_ncu_upconv2_final python tests/prof_upconv.py 2 16 > gpurun_out/c50_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/c50_ncu.log
