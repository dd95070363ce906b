#!/bin/bash
# after the running-column-sum / two-stage pool epilogues: the whole GPU suite, smoke, launch list and full captures of the
# haloed-patch kernel's instantiations inside one training step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c43_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/c43_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c43_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c43_smoke.log
timeout 300 python tests/prof_step.py 3 16 > gpurun_out/c43_plain_step.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_batch16_v4.csv python tests/prof_step.py 3 16 > gpurun_out/c43_ncu_step.log 2>&1
echo "launch list rc=$?"
timeout 300 python tests/prof_step.py 1 16 > gpurun_out/c43_plain_step1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm2_kernel -c 16 -f -o gpurun_out/r2_ncu_igemm2_pairs_step python tests/prof_step.py 1 16 > gpurun_out/c43_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/c43_ncu_full.log
ls -la gpurun_out/*.ncu-rep
