#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/prof_upconv.py 1 > gpurun_out/c28_up1.log 2>&1; echo "rc=$?"; cat gpurun_out/c28_up1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_pair -s 2 -c 1 -o gpurun_out/c28_ncu_up1_fwd -f python tests/prof_upconv.py 1 > gpurun_out/c28_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c28_ncu.log
