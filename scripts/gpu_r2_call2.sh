# Round 2, call 2: full GPU suite after the ABI-2 refactor (per-call options, fused Siddon entry, pose kernels),
# then bench.py with both configurations and a walk-vs-checked A/B of the Siddon forward.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_both.json 2> gpurun_out/r2_bench_both.err; tail -c 6000 gpurun_out/r2_bench_both.json; tail -5 gpurun_out/r2_bench_both.err
timeout 600 python scripts/siddon_ab.py > gpurun_out/r2_siddon_ab.json 2> gpurun_out/r2_siddon_ab.err; cat gpurun_out/r2_siddon_ab.json; tail -5 gpurun_out/r2_siddon_ab.err
