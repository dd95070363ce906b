#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/c30_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/c30_tests.log
timeout 900 python bench.py > gpurun_out/c30_bench_default.json 2> gpurun_out/c30_bench_default.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/c30_bench_default.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d.get("clocks"))
print("roofline", {k:v for k,v in d["roofline"].items() if k!="families"})
print("inference", d.get("inference"))
print("gpu_baseline", d.get("gpu_baseline"))
print("cpu_baseline", d.get("cpu_baseline"))
PY
