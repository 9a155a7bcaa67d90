#!/bin/bash
V=$PWD/build_probe/v
python scripts/sweep_tiles.py trilinear 0,3 2>&1 | tail -1
for v in "$@"; do XVR_B200_LIB=$V/$v.so python scripts/sweep_tiles.py trilinear 0,3 2>&1 | tail -1; done
