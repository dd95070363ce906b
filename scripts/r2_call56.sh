#!/bin/bash
# (not kept: the 40-kernel report was 77 MB, over the 64 MiB pull limit -- capture fewer kernels per report)
mkdir -p gpurun_out
timeout 300 python tests/prof_step.py 1 16 > gpurun_out/c56_plain_step1.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"igemm_pair_kernel|wgrad_pair_kernel|igemm_kernel" -c 40 -f -o gpurun_out/r2_ncu_pair_kernels_final_step python tests/prof_step.py 1 16 > gpurun_out/c56_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/c56_ncu_full.log; ls -la gpurun_out/r2_ncu_pair_kernels_final_step.ncu-rep
