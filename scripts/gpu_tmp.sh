mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "knn or free_running or batch_consistency" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab35.json 2> gpurun_out/r2/ab35.err
tail -2 gpurun_out/r2/ab35.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2/ab35.json"))
st=d["stages_ms"]
print("filter ffma2", round(d["value"]), round(d["ms_per_step"],3), d["checked"], {k:v for k,v in st.items() if "filter" in k})
PY
