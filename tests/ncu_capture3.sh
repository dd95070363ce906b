#!/bin/bash
# Round-1 final ncu --set full captures of igemm2 after the store-warp / mask-tile / fused-pool / fused-bias work.
# (-k matches the base kernel name only: the launches are picked by their position in one training step.)
set -u
cap() {
  local name=$1 skip=$2
  ncu --set full --clock-control none --import-source on -k regex:igemm2_kernel -s "$skip" -c 1 -o gpurun_out/ncu3_$name -f \
      python tests/prof_step.py 1 16 > gpurun_out/ncu3_$name.log 2>&1
  echo "$name rc=$?"
}
cap conv12c_pool_fwd   0     # 1st igemm2 launch of the step: conv12c + fused 2x2 pool
cap conv12e_dgrad_bias 10    # 11th: first of the backward pass, conv12e dgrad + mask tiles + fused bias sums
