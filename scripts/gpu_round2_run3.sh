set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm_persistent or vn_linear or wave_schedule" 2>&1 | tail -5
echo "=== new tests"
timeout 1500 python -m pytest tests/test_gpu_next.py -q -x 2>&1 | tail -40
echo "=== full gpu suite"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30
echo "=== A/B"
for cfg in "LS_GEMM_VARIANT=1" "LS_GEMM_VARIANT=2"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 > gpurun_out/r2/ab3_$tag.json 2> gpurun_out/r2/ab3_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab3_$tag.json"))
    st=d["stages_ms"]
    print("$cfg", round(d["value"]), round(d["ms_per_step"],3), {k:v for k,v in st.items() if "gemm" in k or "global" in k or "head" in k})
except Exception as e:
    print("$cfg FAILED", e)
PY
done
