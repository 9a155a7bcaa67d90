#!/bin/bash
timeout 300 python -m pytest tests/test_siddon_gpu.py tests/test_golden_gpu.py tests/test_volume_gradient_gpu.py -x -q -m gpu 2>&1 | tail -1
python scripts/sweep_tiles.py siddon:64 3,3 2>&1 | tail -1
