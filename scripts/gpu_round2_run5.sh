set -x
mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40
echo "=== configs"
timeout 900 python scripts/run_configs.py > gpurun_out/r2/configs_a.json 2> gpurun_out/r2/configs_a.err; tail -3 gpurun_out/r2/configs_a.err; cat gpurun_out/r2/configs_a.json
