#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/prof_hbm.py 16 > gpurun_out/c32_hbm.log 2>&1
echo "prof rc=$?"; cat gpurun_out/c32_hbm.log | tail -80
