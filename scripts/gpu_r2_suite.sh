#!/bin/bash
# Whole GPU suite + smoke + the bench line (both configurations).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/r2c_bench_N1.json 2> gpurun_out/r2c_bench_N1.err
tail -3 gpurun_out/r2c_bench_N1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench_N1.json"))
for k, v in (("C2", d), ("C5", d.get("config5_siddon", {}))):
    if v:
        t = v.get("empty_space_trimming", {})
        print(k, "value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f frac %.3f marched %.4f untrimmed ms %.2f" % (v["value"], v["e2e"]["value"], v["ms_per_step"], v["roofline"]["kernel_ms"], v["roofline"]["frac"], t.get("marched_fraction", 0), t.get("ms_per_step_without_trimming", 0)))
PY
