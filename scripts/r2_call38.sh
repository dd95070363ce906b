#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/prof_epi_clocks.py 16 > gpurun_out/c38_clocks.log 2>&1
echo "prof rc=$?"; cat gpurun_out/c38_clocks.log
