#!/bin/bash
python scripts/sweep_tiles.py siddon:64 3,3 2,2 2,3 3,4 4,4 4,5 5,5 2,4 3,5 2>&1 | tail -1
timeout 300 python -m pytest tests/test_siddon_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -1
