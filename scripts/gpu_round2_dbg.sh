set -x
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q 2>&1 | grep -E "passed|failed|Error|FAILED|^tests" | head -20
