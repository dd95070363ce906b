#!/bin/bash
# final validation of HEAD: whole GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c55_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/c55_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c55_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c55_smoke.log
SECONDS=0
timeout 1200 python bench.py > gpurun_out/c55_bench_default.json 2> gpurun_out/c55_bench_default.err
echo "default bench rc=$? in ${SECONDS}s"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c55_bench_reference.json 2> gpurun_out/c55_bench_reference.err
echo "reference arm rc=$? in ${SECONDS}s"; cat gpurun_out/c55_bench_reference.json | cut -c1-600
python - <<PY
import json
d=json.load(open("gpurun_out/c55_bench_default.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["whole_step"], d["inference"]["value"], d["gpu_baseline"]["value"])
PY
