#!/bin/bash
# the driver's N=2 launch of bench.py (both arms) + the 2-rank NCCL correctness test
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c_bench_N2.json 2> gpurun_out/r2c_bench_N2.err
tail -3 gpurun_out/r2c_bench_N2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench_N2.json"))
for k, v in (("C2", d), ("C5", d.get("config5_siddon", {}))):
    if v:
        print(k, "n_gpus", v["n_gpus"], "value %.0f e2e %.0f ms/step %.3f" % (v["value"], v["e2e"]["value"], v["ms_per_step"]))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-260
timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -2
for H in 128 256; do
  timeout 600 python scripts/bench_train.py --graph --height $H 2>&1 | grep ms_per_step | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 H=$H N=1', round(d['ms_per_step'],2), 'ms/step')"
  timeout 600 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29502 scripts/bench_train.py --graph --height $H --shard accumulation 2>&1 | grep ms_per_step | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 accumulation-sharded H=$H N=2', round(d['ms_per_step'],2), 'ms/step')"
done
