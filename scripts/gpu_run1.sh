set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
python -c "import os,torch;print(os.cpu_count(), torch.get_num_threads())"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_first.json 2> gpurun_out/bench_r1_first.err; tail -3 gpurun_out/bench_r1_first.err; cat gpurun_out/bench_r1_first.json
for t in 3,4 5,5 0,0 4,4 2,4 1,4 5,8 0,3; do echo "TILE $t"; XVR_B200_TILE=$t timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['kernel_share_of_step'])
    else: print(l.strip()[:200])
"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_first.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trilinear_fwd -s 2 -c 1 -o gpurun_out/prof_tri_fwd_r1a python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
