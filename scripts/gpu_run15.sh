set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python scripts/bench_train.py 2>&1 | tail -1 | cut -c1-330
timeout 600 python scripts/bench_train.py --height 256 --vol 512 --n-vols 1 2>&1 | tail -1 | cut -c1-330
