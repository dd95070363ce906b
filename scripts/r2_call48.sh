#!/bin/bash
# final state of round 2: whole GPU suite, smoke, the default bench line (training + inference + cuDNN + CPU baselines),
# launch list of one training step, full ncu capture of the re-packed wgrad4, per-layer table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c48_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/c48_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c48_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c48_smoke.log
SECONDS=0
timeout 1200 python bench.py > gpurun_out/c48_bench_default.json 2> gpurun_out/c48_bench_default.err
echo "default bench rc=$? in ${SECONDS}s"
python - <<PY
import json
d=json.load(open("gpurun_out/c48_bench_default.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["whole_step"])
print({k:(v if not isinstance(v,dict) else "...") for k,v in d.get("inference",{}).items()})
print(d.get("gpu_baseline"))
print(d.get("cpu_baseline"))
PY
timeout 300 python tests/prof_step.py 3 16 > gpurun_out/c48_plain_step.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_batch16_v5.csv python tests/prof_step.py 3 16 > gpurun_out/c48_ncu_step.log 2>&1
echo "launch list rc=$?"
timeout 300 python tests/prof_step.py 1 16 > gpurun_out/c48_plain_step1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad4_kernel -c 3 -f -o gpurun_out/r2_ncu_wgrad4_ten_slots python tests/prof_step.py 1 16 > gpurun_out/c48_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/c48_ncu_full.log
timeout 900 python tests/prof_layers.py 16 > gpurun_out/r2_layer_table_batch16_v3_final.log 2>&1; echo "layers rc=$?"; tail -4 gpurun_out/r2_layer_table_batch16_v3_final.log
