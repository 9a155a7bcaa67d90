mkdir -p gpurun_out
timeout 200 python scripts/bench_register.py 512 | cut -c1-600
timeout 200 python scripts/bench_register.py 512 --fused-similarity | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:volume_grad_brick -c 1 -f -o gpurun_out/r2_prof_volgrad python scripts/prof_volgrad.py trilinear 8 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:siddon_volume_grad -c 1 -f -o gpurun_out/r2_prof_siddon_volgrad python scripts/prof_volgrad.py siddon 2 2>&1 | tail -2
