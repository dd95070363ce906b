#!/bin/bash
# compute-sanitizer memcheck over every kernel family incl. the round-2 paths (pairs, image-spanning boxes, wgrad4 packing)
mkdir -p gpurun_out
timeout 300 python tests/sanitize_smoke.py > gpurun_out/c52_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/c52_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tests/sanitize_smoke.py > gpurun_out/c52_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/c52_memcheck.log
