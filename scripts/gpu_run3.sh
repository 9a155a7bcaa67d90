set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
for g in tex ldg; do for t in 0,3 0,0 1,4 3,4; do echo "GATHER $g TILE $t"; XVR_B200_GATHER=$g XVR_B200_TILE=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['kernel_share_of_step'], d['e2e']['value'])
    else: print(l.strip()[:300])
"; done; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trilinear_fwd -s 2 -c 1 -o gpurun_out/prof_tri_fwd_r1b python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1
