#!/bin/bash
timeout 900 python -m pytest tests/test_trilinear_gpu.py tests/test_registration_gpu.py tests/test_regsim_gpu.py tests/test_golden_gpu.py tests/test_staged_gpu.py -q -m gpu 2>&1 | tail -3
python scripts/sweep_tiles.py trilinear 0,3 0,3@1 0,3@2 0,3@3 0,2@1 0,2@2 0,1@2 0,1@3 2>&1 | tail -1
timeout 200 python scripts/bench_register.py 512 | cut -c1-700
