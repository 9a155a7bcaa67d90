# First GPU call of round 2: run what round 1 wrote after its GPU budget was spent, then time it.
#   gpurun --timeout 600 -- 'bash scripts/gpu_round2_first.sh'
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zzz_unrun_gpu.py tests/test_regsim_gpu.py -m gpu -q -rxX --runxfail 2>&1 | tail -40
# the staged-brick kernel waits on an mbarrier: first run under its own short timeout
XVR_B200_RUN_UNVALIDATED=1 timeout 120 python -m pytest tests/test_zzz_unrun_gpu.py -m gpu -q -k staged --runxfail 2>&1 | tail -40
for S in 1 2; do
  XVR_B200_STAGED=$S timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_staged$S.json 2> gpurun_out/bench_staged$S.err; cat gpurun_out/bench_staged$S.json
done
timeout 200 python scripts/bench_register.py 512 > gpurun_out/register_unfused.json 2> gpurun_out/register_unfused.err
timeout 200 python scripts/bench_register.py 512 --fused-similarity > gpurun_out/register_fused.json 2> gpurun_out/register_fused.err
cat gpurun_out/register_unfused.json gpurun_out/register_fused.json
