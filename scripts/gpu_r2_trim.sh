mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trilinear_gpu.py tests/test_zz_full_size_gpu.py tests/test_golden_gpu.py tests/test_trainer_gpu.py tests/test_registration_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --config trilinear --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_trim.json 2> gpurun_out/r2_bench_trim.err; tail -3 gpurun_out/r2_bench_trim.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_trim.json').read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), "kernel_ms", round(d["roofline"]["kernel_ms"],2), "frac", round(d["roofline"]["frac"],3))
print(d["empty_space_trimming"])
PY
