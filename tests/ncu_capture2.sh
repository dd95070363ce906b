#!/bin/bash
# Round-1 (second half) ncu evidence: launch list of the batch-16 training step and --set full captures of the kernels
# that changed: fused conv12e+head, linear loss, conv11c fwd / wgrad, pool backward with bias sums, head backward,
# weight-map kernels.  Run on a B200 box via gpurun; everything lands in gpurun_out/.
set -u
python tests/prof_step.py 2 16 > gpurun_out/ncu2_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches7.csv python tests/prof_step.py 2 16 > gpurun_out/ncu2_list.log 2>&1
echo "launch list rc=$?"
cap() {  # name regex skip script...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s "$skip" -c 1 -o gpurun_out/ncu2_$name -f \
      python "$@" > gpurun_out/ncu2_$name.log 2>&1
  echo "$name rc=$?"
}
cap conv12e_head   'igemm2_kernel<64, 1, true>'  0 tests/prof_step.py 1 16
cap conv12c_dgrad  'igemm2_kernel<64, 1, false>' 2 tests/prof_step.py 1 16
cap loss_linear    'loss_linear_kernel'          0 tests/prof_step.py 1 16
cap conv1_fwd      'conv1_fwd_kernel'            0 tests/prof_step.py 1 16
cap conv1_wgrad    'conv1_wgrad_kernel'          0 tests/prof_step.py 1 16
cap pool_bwd_bias  'pool_bwd_kernel'             3 tests/prof_step.py 1 16
cap head_bwd       'head_bwd_kernel'             0 tests/prof_step.py 1 16
python tests/prof_weightmap.py 16 1 > gpurun_out/ncu2_wm_plain.log 2>&1
cap wm_rows        'wm_rows_kernel'              1 tests/prof_weightmap.py 16 1
cap wm_columns     'wm_columns_kernel'           1 tests/prof_weightmap.py 16 1
cap wm_merge       'wm_merge_kernel'             1 tests/prof_weightmap.py 16 1
