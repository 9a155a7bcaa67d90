# Round profile (round 2): bench line (both configurations) + reference arm, the ncu launch list of the same command,
# ncu --set full of the dominant kernels.   gpurun --timeout 2400 -- 'bash scripts/gpu_profile.sh r2'
TAG=${1:-r2}
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_N1.json 2> gpurun_out/${TAG}_bench_N1.err
tail -c 300 gpurun_out/${TAG}_bench_N1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
cat gpurun_out/${TAG}_bench_ref.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trilinear_fwd_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_trilinear_fwd python bench.py --config trilinear --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:siddon_fwd_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_siddon_fwd python bench.py --config siddon --batch 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_siddon.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_siddon.log
XVR_B200_STAGED=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:trilinear_fwd_staged -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_staged python bench.py --config trilinear --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_staged.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_staged.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:volume_grad_brick -c 1 -f -o gpurun_out/${TAG}_prof_volgrad python scripts/prof_volgrad.py > gpurun_out/${TAG}_ncu_volgrad.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_volgrad.log
ls -la gpurun_out/*.ncu-rep
