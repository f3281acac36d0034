mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -k "encode or shape_prior or c2 or c3 or batch_consistency or free_running or teacher" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab23.json 2> gpurun_out/r2/ab23.err
tail -2 gpurun_out/r2/ab23.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2/ab23.json"))
st=d["stages_ms"]
print("gather+rowmean", round(d["value"]), round(d["ms_per_step"],3), d["checked"], {k:v for k,v in st.items() if "gather" in k or "global" in k})
PY
