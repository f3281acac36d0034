set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "vn_linear or gemm_persistent or wave_schedule or teacher" 2>&1 | tail -5
for mk in 0 128; do
env LS_GEMM_V3_MIN_K=$mk timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab12_mk$mk.json 2> gpurun_out/r2/ab12_mk$mk.err
tail -3 gpurun_out/r2/ab12_mk$mk.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab12_mk$mk.json"))
    st=d["stages_ms"]
    print("min_k $mk", round(d["value"]), round(d["ms_per_step"],3), {k:v for k,v in st.items() if "global" in k or "gemm" in k or "head" in k})
except Exception as e:
    print("FAILED", e)
PY
done
