#!/bin/bash
# compute-sanitizer memcheck over the tcgen05 attention kernels (GPT-2 and the stage-1 relative-position mode, incl. the
# shift / un-shift kernels and the ragged-batch decode step).  gpurun --timeout 1500 -- bash scripts/sanitize_attn.sh
set -u
mkdir -p gpurun_out
timeout 1300 compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 120 \
  python -m pytest tests/test_gpt2_gpu.py tests/test_stage1_gpu.py tests/test_decode_gpu.py -x -q \
  -k "attention_fwd_bwd_vs_torch or tcgen05_equals_mma_sync or ragged_batch" -p no:cacheprovider > gpurun_out/sanitize_attn_memcheck.log 2>&1
rc=$?
echo "compute-sanitizer memcheck (attention): rc=$rc  $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_attn_memcheck.log | tail -1)  |  $(tail -1 gpurun_out/sanitize_attn_memcheck.log)"
grep -E "Invalid|out of bounds|misaligned" gpurun_out/sanitize_attn_memcheck.log | sort | uniq -c | head -10
