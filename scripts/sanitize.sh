#!/bin/bash
# compute-sanitizer pass over the tcgen05 / TMA kernels (memcheck + racecheck per kernel).  Run on
# the GPU box through gpurun; writes gpurun_out/sanitize_<tool>.log and a one-line verdict per tool.  The kernels
# carry five mbarrier families and (GEMM, CG = 2) a cross-CTA protocol; the watchdog in mbar_wait only catches hangs,
# this catches out-of-bounds shared / global / tensor-memory accesses and unsynchronised shared-memory hazards.
#   gpurun --timeout 1500 -- bash scripts/sanitize.sh
set -u
mkdir -p gpurun_out
SEL=${SEL:-"gemm or favor"}
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 --launch-timeout 120 \
    python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitize_$tool.log 2>&1
  rc=$?
  echo "compute-sanitizer $tool: rc=$rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)  |  $(tail -1 gpurun_out/sanitize_$tool.log)"
done
