# multi-GPU call (8 GPUs of one box): the 2-rank NCCL correctness test, C4 strong scaling at 1/2/4/8, bench.py at 2 and 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5
for H in 128 256; do
  timeout 300 python scripts/bench_train.py --graph --height $H 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 H=$H N=1', round(d['ms_per_step'],2), 'ms/step', d['last_log'])"
  for N in 2 4 8; do
    timeout 600 $TR --nproc-per-node $N --master-port $((29500+N)) scripts/bench_train.py --graph --height $H 2>&1 | grep ms_per_step | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 H=$H N=$N', round(d['ms_per_step'],2), 'ms/step', d['last_log'])"
  done
done
timeout 600 $TR --nproc-per-node 8 --master-port 29600 scripts/bench_train.py --graph --profile 2>&1 | grep -E "Name|void|nccl|Self CUDA time" | cut -c1-72,150-260 | head -22
timeout 900 $TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_N2.json 2> gpurun_out/r2_bench_N2.err; tail -c 400 gpurun_out/r2_bench_N2.json; tail -2 gpurun_out/r2_bench_N2.err
timeout 900 $TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_N8.json 2> gpurun_out/r2_bench_N8.err; python - <<'PY'
import json
for n in (2, 8):
    try:
        d=json.loads(open(f'gpurun_out/r2_bench_N{n}.json').read().strip().splitlines()[-1])
        print("bench N", n, round(d["value"]), "DRR/s e2e", round(d["e2e"]["value"]), "| siddon", round(d["config5_siddon"]["value"], 1), "DRR/s")
    except Exception as e: print("bench N", n, "failed", e)
PY
timeout 600 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 --impl reference --steps 5 --warmup 2 2>/dev/null | tail -1 | cut -c1-900
