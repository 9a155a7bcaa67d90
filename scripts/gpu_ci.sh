# What the round-end driver does, in one call:  gpurun -- 'bash scripts/gpu_ci.sh'
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/bench_N1.json 2> gpurun_out/bench_N1.err; cat gpurun_out/bench_N1.json
