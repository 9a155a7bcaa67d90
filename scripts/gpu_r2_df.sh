#!/bin/bash
# distance-field walk: trimming tests for every brick size, then timing
V=$PWD/build_probe/v
for v in b8 b4 b16; do
  echo "== $v"
  XVR_B200_LIB=$V/$v.so timeout 300 python -m pytest tests/test_trilinear_gpu.py tests/test_siddon_gpu.py -q -m gpu -k "trim" -s 2>&1 | grep -v "^$" | tail -8
  XVR_B200_LIB=$V/$v.so python scripts/sweep_tiles.py trilinear 0,3 2>&1 | tail -1
  XVR_B200_LIB=$V/$v.so python scripts/sweep_tiles.py siddon:64 3,3 2>&1 | tail -1
done
