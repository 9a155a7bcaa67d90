#!/bin/bash
# Siddon empty-space trimming: bit-identity tests, then both bench configurations.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_siddon_gpu.py tests/test_trilinear_gpu.py tests/test_golden_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -25
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_sidtrim.json 2> gpurun_out/r2_bench_sidtrim.err
tail -5 gpurun_out/r2_bench_sidtrim.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_sidtrim.json"))
for k, v in (("C2", d), ("C5", d.get("config5_siddon", {}))):
    if v:
        print(k, "value %.0f e2e %.0f ms/step %.2f frac %.3f" % (v["value"], v["e2e"]["value"], v["ms_per_step"], v["roofline"]["frac"]))
        t = v.get("empty_space_trimming", {})
        print({kk: t[kk] for kk in t if kk != "what"})
PY
