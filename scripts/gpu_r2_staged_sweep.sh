# staged-brick kernel: tests with the default build, then timing + shared-memory share of tile/ring variants
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_staged_gpu.py -m gpu -x -q 2>&1 | tail -8
XVR_B200_LIB=build_probe/libxvr_E.so timeout 300 python scripts/staged_stats.py 2>&1 | tail -1
for V in F G H; do
  export XVR_B200_LIB=build_probe/libxvr_$V.so
  XVR_B200_STAGED=1 timeout 300 python bench.py --config trilinear --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_staged_$V.json 2> gpurun_out/r2_staged_$V.err
  python - "$V" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2_staged_{v}.json').read().strip().splitlines()[-1])
    print("VARIANT", v or "A", "DRR/s", round(d["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"], 2), d["roofline"]["entry_point"])
except Exception as e: print("VARIANT", v, "no bench line", e)
PY
  timeout 300 python scripts/staged_stats.py 2>&1 | tail -1
done
