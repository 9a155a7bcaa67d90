set -x
for v in "3 4" "4 4" "4 2" "3 2" "3 8" "2 4" "4 1"; do set -- $v
  XVR_B200_NVCC_FLAGS="-DXVR_TRI_MIN_CTAS=$1 -DXVR_TRI_UNROLL=$2" python -c "from xvr_b200 import _build; _build.build(force=True)" > /dev/null 2>&1
  grep -A3 "trilinear_fwd_kernelILb1ELb0ELb1" xvr_b200/build/ptxas.log | grep -E "registers|spill" | tr '\n' ' '
  echo "VARIANT min_ctas=$1 unroll=$2"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])
    else: print(l.strip()[:300])
"
done
