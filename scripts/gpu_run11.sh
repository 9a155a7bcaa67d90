set -x
timeout 900 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x 2>&1 | tail -25
timeout 600 python scripts/bench_train.py 2>&1 | tail -2
timeout 600 python scripts/bench_train.py --labels 2>&1 | tail -2
timeout 600 python scripts/bench_train.py --height 256 --vol 512 --n-vols 1 2>&1 | tail -2
