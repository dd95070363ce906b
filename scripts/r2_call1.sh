#!/bin/bash
# round-2 diagnostics: same-box baseline numbers + full ncu captures of the two largest kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt
for c in "wg42 0" "wg52 0" "wg41 0" "wg41e 0" "fw42 0" "fw52 0" "fw41e 0"; do
  python tests/prof_kernels.py $c 5 >> gpurun_out/c1_prof.log 2>&1 || echo "FAIL $c" >> gpurun_out/c1_prof.log
done
python bench.py --impl cudnn --steps 5 --warmup 3 > gpurun_out/c1_cudnn.json 2> gpurun_out/c1_cudnn.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
python tests/prof_kernels.py wg42 0 3 > gpurun_out/c1_plain_wg42.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_wgrad256_conv42c python tests/prof_kernels.py wg42 0 3 > gpurun_out/c1_ncu1.log 2>&1
python tests/prof_kernels.py fw42 0 3 > gpurun_out/c1_plain_fw42.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_igemm256_conv42c python tests/prof_kernels.py fw42 0 3 > gpurun_out/c1_ncu2.log 2>&1
python tests/prof_kernels.py wg52 0 3 > gpurun_out/c1_plain_wg52.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_wgrad256_conv52c python tests/prof_kernels.py wg52 0 3 > gpurun_out/c1_ncu3.log 2>&1
ls -la gpurun_out
