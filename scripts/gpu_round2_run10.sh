set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm or vn_linear or wave_schedule or free_running or teacher or sdf" 2>&1 | tail -5
env timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab10.json 2> gpurun_out/r2/ab10.err
tail -3 gpurun_out/r2/ab10.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab10.json"))
    st=d["stages_ms"]
    print("epi2", round(d["value"]), round(d["ms_per_step"],3), {k:v for k,v in st.items() if "global" in k or "gemm" in k or "head" in k})
except Exception as e:
    print("FAILED", e)
PY
timeout 300 python bench.py --workload sdf --no-cpu-baseline 2>&1 | cut -c1-400
