set -x
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -30
echo "=== A/B"
for cfg in "LS_GEMM_VARIANT=2"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 > gpurun_out/r2/ab4_$tag.json 2> gpurun_out/r2/ab4_$tag.err
  tail -3 gpurun_out/r2/ab4_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab4_$tag.json"))
    st=d["stages_ms"]
    print("$cfg", round(d["value"]), round(d["ms_per_step"],3), {k:v for k,v in st.items() if "gemm" in k or "global" in k or "head" in k})
except Exception as e:
    print("$cfg FAILED", e)
PY
done
