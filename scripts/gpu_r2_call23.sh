mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rs -s -k "not staged" 2>&1 | grep -E "arbiter|passed|failed|FAILED|Error|assert|SKIP" | head -40
