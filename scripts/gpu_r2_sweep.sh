#!/bin/bash
V=$PWD/build_probe/v
XVR_B200_LIB=$V/base.so python scripts/sweep_tiles.py trilinear 0,3 1,3 2,3 2,4 3,3 3,5 0,2 0,1 1,2 2>&1 | tail -1
XVR_B200_LIB=$V/cta4.so python scripts/sweep_tiles.py trilinear 0,3 2,3 0,2 2>&1 | tail -1
XVR_B200_LIB=$V/cta2.so python scripts/sweep_tiles.py trilinear 0,3 2,3 2>&1 | tail -1
XVR_B200_LIB=$V/un2.so python scripts/sweep_tiles.py trilinear 0,3 2,3 2>&1 | tail -1
XVR_B200_LIB=$V/un8.so python scripts/sweep_tiles.py trilinear 0,3 2,3 2>&1 | tail -1
XVR_B200_LIB=$V/base.so python scripts/sweep_tiles.py siddon:64 0,3 1,3 2,3 2,4 3,3 0,2 2>&1 | tail -1
