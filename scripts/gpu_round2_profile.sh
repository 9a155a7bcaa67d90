# Round 2, after scripts/gpu_round2_first.sh is green: ncu captures of the kernels written at the end of round 1.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_profile.sh'
set -x
mkdir -p gpurun_out
for S in 1 2; do
  XVR_B200_STAGED=$S timeout 400 ncu --set full --clock-control none --import-source on -k regex:trilinear_fwd_staged \
    -s 3 -c 1 -f -o gpurun_out/prof_trilinear_staged$S python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_staged$S.log 2>&1
  tail -3 gpurun_out/ncu_staged$S.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:regsim -c 3 -f -o gpurun_out/prof_regsim \
  python scripts/bench_register.py 512 --fused-similarity > gpurun_out/ncu_regsim.log 2>&1
tail -3 gpurun_out/ncu_regsim.log
