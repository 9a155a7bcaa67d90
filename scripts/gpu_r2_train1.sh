mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_volume_gradient_gpu.py -m gpu -x -q 2>&1 | tail -15
for B in 116 15; do
  timeout 300 python scripts/bench_train.py --graph --batch $B --profile > gpurun_out/r2_train_b$B.txt 2>&1; head -c 1200 gpurun_out/r2_train_b$B.txt | head -3; grep -E "Name|void|nccl|Memcpy|Memset|kernel" gpurun_out/r2_train_b$B.txt | cut -c1-72,150-260 | head -26
done
timeout 300 python scripts/bench_train.py --graph --batch 15 --bf16 --channels-last 2>&1 | tail -1 | cut -c1-400
timeout 300 python scripts/bench_train.py --graph --batch 116 --bf16 --channels-last 2>&1 | tail -1 | cut -c1-400
