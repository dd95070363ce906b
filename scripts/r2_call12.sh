#!/bin/bash
# 2 GPUs: DP correctness in the three all-reduce modes, then bench at 16 tiles / GPU in each
mkdir -p gpurun_out
for m in symm peer nccl; do
  UB_DP_ALLREDUCE=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/gpu_dp_check.py > gpurun_out/c12_dpcheck_$m.log 2>&1
  echo "== dp_check $m rc=$?"; grep -E "PASS|FAIL|multicast|Error|error" gpurun_out/c12_dpcheck_$m.log | head -6
done
run() {
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 10 --warmup 3 --batch-per-gpu 16 --no-inference > gpurun_out/c12_${name}.json 2> gpurun_out/c12_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c12_${name}.json"))
    b=d.get("scaling_breakdown") or {}
    print("${name}", round(d["value"],1), "tiles/s", round(d["ms_per_step"],3), "ms | eager", round(b.get("compute_only_eager_ms",0),3), "graph", round(b.get("compute_only_graph_ms",0),3), "exposed", round(b.get("exposed_allreduce_ms",0),3), "e2e", round(d["e2e"]["value"],1), "|", b.get("allreduce"))
except Exception as e:
    print("${name} failed", e)
PY
  tail -2 gpurun_out/c12_${name}.err | cut -c1-300
}
run symm UB_DP_ALLREDUCE=symm
run peer UB_DP_ALLREDUCE=peer
run nccl UB_DP_ALLREDUCE=nccl
run symm_eager UB_DP_ALLREDUCE=symm UB_DP_GRAPH=0
