#!/bin/bash
# A/B of tuning variants built by scripts/build_variant.py: trimming bit-identity tests + both bench configurations each.
mkdir -p gpurun_out
for v in "$@"; do
  export XVR_B200_LIB=$PWD/build_probe/v/$v.so
  echo "=== $v"
  timeout 300 python -m pytest tests/test_trilinear_gpu.py tests/test_siddon_gpu.py -x -q -m gpu -k "trim" 2>&1 | tail -1
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  tail -3 gpurun_out/ab_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$v.json"))
for k, v in (("C2", d), ("C5", d.get("config5_siddon", {}))):
    if v:
        t = v.get("empty_space_trimming", {})
        print(k, "value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f marched %.4f untrimmed ms %.2f" % (v["value"], v["e2e"]["value"], v["ms_per_step"], v["roofline"]["kernel_ms"], t.get("marched_fraction", 0), t.get("ms_per_step_without_trimming", 0)))
PY
done
