#!/bin/bash
N=8
mkdir -p gpurun_out
run() {
  name=$1; shift; extra=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 $extra > gpurun_out/c14_n8_${name}.json 2> gpurun_out/c14_n8_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c14_n8_${name}.json"))
    b=d.get("scaling_breakdown") or {}
    print("${name}", round(d["value"],1), "tiles/s", round(d["ms_per_step"],3), "ms | eager", round(b.get("compute_only_eager_ms",0),3), "graph", round(b.get("compute_only_graph_ms",0),3), "exposed", round(b.get("exposed_allreduce_ms",0),3), "e2e", round(d["e2e"]["value"],1), "|", b.get("allreduce"))
    i=d.get("inference")
    if i: print("  inference", round(i["value"],1), "Mpix/s device,", round(i["e2e"]["value"],1), "e2e")
except Exception as e:
    print("${name} failed", e)
PY
  tail -2 gpurun_out/c14_n8_${name}.err | cut -c1-300
}
run symm "" UB_DP_ALLREDUCE=symm
run peer "--no-inference" UB_DP_ALLREDUCE=peer
run nccl "--no-inference" UB_DP_ALLREDUCE=nccl
