set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 600 python scripts/bench_register.py 512 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
