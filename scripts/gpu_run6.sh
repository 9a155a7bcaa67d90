set -x
timeout 1500 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -40
