set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:siddon_fwd -s 2 -c 1 -o gpurun_out/prof_siddon_fwd python scripts/bench_kernels.py --only siddon --quick > gpurun_out/ncu_siddon.log 2>&1
tail -3 gpurun_out/ncu_siddon.log
