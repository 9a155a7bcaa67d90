#!/bin/bash
# Whole GPU suite + smoke + the bench line (both configurations, CPU baseline) + the reference arm + config 3.
mkdir -p gpurun_out
T=${TAG:-r2c}
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_N1.json 2> gpurun_out/${T}_bench_N1.err
tail -3 gpurun_out/${T}_bench_N1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
cut -c1-200 gpurun_out/${T}_bench_ref.json
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_N1.json"))
for k, v in (("C2", d), ("C5", d.get("config5_siddon", {}))):
    if v:
        t = v.get("empty_space_trimming", {})
        print(k, "value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f frac %.3f marched %.4f untrimmed ms %.2f" % (v["value"], v["e2e"]["value"], v["ms_per_step"], v["roofline"]["kernel_ms"], v["roofline"]["frac"], t.get("marched_fraction", 0), t.get("ms_per_step_without_trimming", 0)))
print("cpu_baseline", d.get("cpu_baseline", {}).get("value"))
PY
timeout 200 python scripts/bench_register.py 512 | cut -c1-400
timeout 200 python scripts/bench_register.py 512 --unfused-similarity | cut -c1-400
