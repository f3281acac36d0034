mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "teacher or free_running or batch_consistency or wave_schedule or encode" 2>&1 | tail -3
for so in 0 1; do
env LS_EDGE_STAGE_OUT=$so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab21_$so.json 2> gpurun_out/r2/ab21_$so.err
tail -2 gpurun_out/r2/ab21_$so.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab21_$so.json"))
    st=d["stages_ms"]
    print("stage_out $so", round(d["value"]), round(d["ms_per_step"],3), d["checked"], {k:v for k,v in st.items() if "edgeconv" in k})
except Exception as e:
    print("FAILED", e)
PY
done
