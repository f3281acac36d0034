mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -k "teacher or free_running or batch_consistency or encode or knn_tc_equals or c2 or c3" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab34.json 2> gpurun_out/r2/ab34.err
tail -2 gpurun_out/r2/ab34.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2/ab34.json"))
st=d["stages_ms"]
print("deep 128 regs", round(d["value"]), round(d["ms_per_step"],3), d["checked"], {k:v for k,v in st.items() if "edgeconv" in k})
PY
