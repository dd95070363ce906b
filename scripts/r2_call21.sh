#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c21_n$N.json 2> gpurun_out/c21_n$N.err
echo "== n$N rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/c21_n$N.json"))
b=d.get("scaling_breakdown") or {}
print(round(d["value"],1), "tiles/s", round(d["ms_per_step"],3), "ms | eager", round(b.get("compute_only_eager_ms",0),3), "graph", round(b.get("compute_only_graph_ms",0),3), "exposed", round(b.get("exposed_allreduce_ms",0),3), "e2e", round(d["e2e"]["value"],1), "|", b.get("allreduce"))
i=d.get("inference"); print("inference", round(i["value"],1), round(i["e2e"]["value"],1))
PY
tail -3 gpurun_out/c21_n$N.err | cut -c1-300
