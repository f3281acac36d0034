mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -k "encode or shape_prior or c2 or batch_consistency or free_running" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab22.json 2> gpurun_out/r2/ab22.err
tail -2 gpurun_out/r2/ab22.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2/ab22.json"))
st=d["stages_ms"]
print("normalize", round(d["value"]), round(d["ms_per_step"],3), d["checked"], st["normalize"], st["fps"], st["head"])
PY
