mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "teacher or free_running or wave_schedule or batch_consistency or encode" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab16.json 2> gpurun_out/r2/ab16.err
tail -2 gpurun_out/r2/ab16.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab16.json"))
    st=d["stages_ms"]
    print("bias+vnact", round(d["value"]), round(d["ms_per_step"],3), d["checked"], {k:v for k,v in st.items() if "global" in k})
except Exception as e:
    print("FAILED", e)
PY
