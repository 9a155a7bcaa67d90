#!/bin/bash
# compute-sanitizer memcheck over the GPU suite (all but the full-size cases): odd volume shapes, label bricks, fused
# label channels, trimming on both renderers, similarity, pose, registration and training kernels.
mkdir -p gpurun_out
timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 \
  python -m pytest tests -q -m gpu --deselect tests/test_zz_full_size_gpu.py --deselect tests/test_multi_gpu.py \
  > gpurun_out/sanitize.log 2>&1
echo "exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize.log | head -20
tail -3 gpurun_out/sanitize.log
