set -x
mkdir -p gpurun_out
./build_probe/probe_tex 2>&1 | tee gpurun_out/probe_tex.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
