set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
for f in 1 0; do echo "FUSED $f"; XVR_B200_FUSED=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['kernel_share_of_step'], d['e2e']['value'], d['clocks'])
    else: print(l.strip()[:300])
"; done
