set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "vn_linear or gemm_persistent" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "wave_schedule or free_running or teacher or sdf" 2>&1 | tail -8
for v in 2 3; do
env LS_GEMM_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab11_v$v.json 2> gpurun_out/r2/ab11_v$v.err
tail -3 gpurun_out/r2/ab11_v$v.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab11_v$v.json"))
    st=d["stages_ms"]
    print("variant $v", round(d["value"]), round(d["ms_per_step"],3), {k:v for k,v in st.items() if "global" in k or "gemm" in k or "head" in k})
except Exception as e:
    print("FAILED", e)
PY
env LS_GEMM_VARIANT=$v timeout 300 python bench.py --workload sdf --no-cpu-baseline 2>&1 | cut -c1-330
done
