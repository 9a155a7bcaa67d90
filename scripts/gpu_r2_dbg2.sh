mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29711 tests/_nccl_worker.py 2>&1 | grep -v "^W1017\|OMP_NUM\|^\*\*\*" | tail -25
python - <<'PY'
import subprocess, sys, json
for b in (15, 116):
    out = subprocess.run([sys.executable, "scripts/bench_train.py", "--graph", "--batch", str(b), "--steps", "16", "--log-every", "1"], capture_output=True, text=True).stdout
    print(b, out.strip().splitlines()[-1][-330:])
PY
