#!/bin/bash
timeout 300 python -m pytest tests/test_siddon_gpu.py -x -q -m gpu -k "random_sparse" 2>&1 | tail -15
