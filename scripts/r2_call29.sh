#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/c29_kernels.log 2>&1
echo "kernels rc=$?"; tail -4 gpurun_out/c29_kernels.log
timeout 600 python tests/prof_rings.py 16 > gpurun_out/c29_rings.log 2>&1
echo "prof rc=$?"; cat gpurun_out/c29_rings.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-inference --no-gpu-baseline > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/c29_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
for k,v in d["roofline"]["families"].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
