#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c15_pytest.log 2>&1; tail -6 gpurun_out/c15_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c15_n2.json 2> gpurun_out/c15_n2.err
echo "== n2 rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/c15_n2.json"))
b=d.get("scaling_breakdown") or {}
print(round(d["value"],1), "tiles/s", round(d["ms_per_step"],3), "ms | eager", round(b.get("compute_only_eager_ms",0),3), "graph", round(b.get("compute_only_graph_ms",0),3), "exposed", round(b.get("exposed_allreduce_ms",0),3), "e2e", round(d["e2e"]["value"],1), "|", b.get("allreduce"))
i=d.get("inference"); print("inference", round(i["value"],1), round(i["e2e"]["value"],1))
PY
tail -3 gpurun_out/c15_n2.err | cut -c1-300
