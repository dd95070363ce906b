#!/bin/bash
# full ncu captures of the first ten CTA-pair launches of a training step (forward: conv31c ... up-convolution 3), final kernels
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source off -k regex:"igemm_pair_kernel" -c 10 -f -o gpurun_out/r2_ncu_pair_fwd_final python tests/prof_step.py 1 16 > gpurun_out/c57_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/c57_ncu_full.log; ls -la gpurun_out/r2_ncu_pair_fwd_final.ncu-rep
