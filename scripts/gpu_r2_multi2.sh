# multi-GPU call 2: NCCL correctness incl. accumulation sharding, C4 with the accumulation window dealt out to the ranks
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6
for H in 128 256; do
  for N in 2 4 8; do
    timeout 600 $TR --nproc-per-node $N --master-port $((29500+N)) scripts/bench_train.py --graph --height $H --shard accumulation 2>&1 | grep ms_per_step | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 accumulation-sharded H=$H N=$N', round(d['ms_per_step'],2), 'ms/step', {k: round(v, 3) for k, v in d['last_log'].items() if k in ('loss','mncc','kept')})"
  done
done
