# full GPU validation: test suite, smoke, registration timing
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 200 python scripts/bench_register.py 512 | cut -c1-500
