#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c20_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/c20_smoke.log
python tests/prof_kernels.py fw42 0 3 > gpurun_out/c20_plain_fw42.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:igemm_pair_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_igemm_pair_conv42c python tests/prof_kernels.py fw42 0 3 > gpurun_out/c20_ncu1.log 2>&1
python tests/prof_kernels.py wg52 0 3 > gpurun_out/c20_plain_wg52.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wgrad_pair_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_wgrad_pair_conv52c python tests/prof_kernels.py wg52 0 3 > gpurun_out/c20_ncu2.log 2>&1
python tests/prof_step.py 3 16 > gpurun_out/c20_plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_batch16_v3.csv python tests/prof_step.py 3 16 > gpurun_out/c20_ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_conv1_wgrad_tc python tests/prof_step.py 2 16 > gpurun_out/c20_ncu3.log 2>&1
cat gpurun_out/c20_plain_*.log
