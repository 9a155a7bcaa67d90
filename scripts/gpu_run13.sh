timeout 900 ncu --set full --clock-control none --import-source on -k regex:volume_grad -s 1 -c 1 -o gpurun_out/prof_volgrad python scripts/prof_volgrad.py > gpurun_out/ncu_volgrad.log 2>&1
tail -2 gpurun_out/ncu_volgrad.log
