#!/bin/bash
V=$PWD/build_probe/v
for v in "$@"; do XVR_B200_LIB=$V/$v.so python scripts/sweep_tiles.py siddon:64 3,3 2>&1 | tail -1; done
