# Launch-bounds / unroll sweep of the trilinear forward kernel (rebuilds on the GPU box, one bench line per variant).
set -x
for v in ${XVR_SWEEP:-"4 4" "4 2" "5 4" "3 4"}; do set -- $v
  XVR_B200_NVCC_FLAGS="-DXVR_TRI_MIN_CTAS=$1 -DXVR_TRI_UNROLL=$2" python -c "from xvr_b200 import _build; _build.build(force=True)" > /dev/null 2>&1
  grep -A3 "trilinear_fwd_kernelILb1ELb0ELb1" xvr_b200/build/ptxas.log | grep -E "registers|spill" | tr '\n' ' '
  echo "VARIANT min_ctas=$1 unroll=$2"
  timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])
"
done
