#!/bin/bash
# ncu --set full captures of one launch per kernel family inside the real batch-16 training step
# (tests/prof_step.py).  Run on a B200 box via gpurun; reports land in gpurun_out/ncu_*.ncu-rep.
set -u
python tests/prof_step.py 1 16 > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c 1 -o gpurun_out/ncu_$1 -f \
      python tests/prof_step.py 1 16 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
cap igemm256_conv42c_fwd   'igemm_kernel<256>'      3
cap igemm2_64_conv12c_fwd  'igemm2_kernel<64, 1>'   0
cap igemm2_128x2_conv22c   'igemm2_kernel<128, 2>'  0
cap wgrad3_conv12e         'wgrad3_kernel'          0
cap wgrad5_conv11e         'wgrad5_kernel'          0
cap wgrad256               'wgrad_kernel<256>'      4
cap pool_fwd_pool1         'pool_fwd_kernel'        0
cap pool_bwd_pool1         'pool_bwd_kernel'        3
cap loss                   'loss_kernel'            0
cap colsum                 'colsum_kernel'          0
cap conv1_fwd              'conv1_fwd_kernel'       0
cap head_bwd               'head_bwd_kernel'        0
cap sgd                    'sgd_kernel'             0
