set -x
timeout 900 python -m pytest tests/test_siddon_gpu.py tests/test_golden_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python scripts/bench_kernels.py --only siddon 2>&1 | tail -8
