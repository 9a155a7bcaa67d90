# staged-brick kernel v3: correctness under a watchful timeout, then timing with the stats counters
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_staged_gpu.py -m gpu -x -q 2>&1 | tail -25
XVR_B200_STAGED=1 timeout 300 python bench.py --config trilinear --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_staged_a.json 2> gpurun_out/r2_staged_a.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_staged_a.json').read().strip().splitlines()[-1])
print("STAGED A:", d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["entry_point"])
PY
tail -3 gpurun_out/r2_staged_a.err
timeout 300 python scripts/staged_stats.py 2>&1 | tail -5
