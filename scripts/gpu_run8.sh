set -x
timeout 900 python -m pytest tests/test_registration_gpu.py -m gpu -q -x 2>&1 | tail -30
timeout 600 python scripts/bench_register.py 512 2>&1 | tail -3
