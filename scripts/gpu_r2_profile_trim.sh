#!/bin/bash
# Round-2 final-state profiles (tag r2c): the ncu launch list of the bench command and ncu --set full of the two forward
# kernels, first launch of the TIMED region (bench.py brackets it with cudaProfilerStart/Stop; the warm-up also holds the
# untrimmed comparison launches).  Numbers printed under ncu are not bench values.
mkdir -p gpurun_out
T=r2c
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:trilinear_fwd_kernel -c 1 -f \
  -o gpurun_out/${T}_prof_trilinear_fwd python bench.py --config trilinear --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
tail -2 gpurun_out/${T}_ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:siddon_fwd_kernel -c 1 -f \
  -o gpurun_out/${T}_prof_siddon_fwd python bench.py --config siddon --batch 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_siddon.log 2>&1
tail -2 gpurun_out/${T}_ncu_siddon.log
ls -la gpurun_out/${T}*.ncu-rep
