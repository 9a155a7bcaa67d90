set -x
timeout 900 python -m pytest tests/test_trilinear_gpu.py -m gpu -q -x -k "volume_gradient" 2>&1 | tail -5
timeout 600 python scripts/bench_kernels.py --only trilinear 2>&1 | grep -i "dvolume"
