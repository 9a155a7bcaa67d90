# ncu --set full captures of the Siddon forward and the volume-gradient kernels -> gpurun_out/
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:siddon_fwd -s 2 -c 1 -f -o gpurun_out/prof_siddon_fwd python scripts/bench_kernels.py --only siddon --quick > gpurun_out/ncu_siddon.log 2>&1
tail -3 gpurun_out/ncu_siddon.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:volume_grad -s 1 -c 1 -f -o gpurun_out/prof_volgrad python scripts/prof_volgrad.py > gpurun_out/ncu_volgrad.log 2>&1
tail -3 gpurun_out/ncu_volgrad.log
