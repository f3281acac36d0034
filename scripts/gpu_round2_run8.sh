set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm_persistent or vn_linear or wave_schedule or free_running or teacher" 2>&1 | tail -5
env timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab8.json 2> gpurun_out/r2/ab8.err
tail -3 gpurun_out/r2/ab8.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab8.json"))
    st=d["stages_ms"]
    print("tn96", round(d["value"]), round(d["ms_per_step"],3), {k:v for k,v in st.items() if "global" in k or "gemm" in k or "head" in k})
except Exception as e:
    print("FAILED", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc2 -s 8 -c 6 -o gpurun_out/r2/prof_gemm2c python scripts/one_forward.py 1 > gpurun_out/r2/prof_gemm2c.log 2>&1
ls -la gpurun_out/r2/*.ncu-rep
