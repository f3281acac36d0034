set -x
timeout 900 python -m pytest tests/test_gpu_next.py -q -s -k "sdf_backward or generator3d" 2>&1 | grep -E "sdf backward|passed|failed|Error|stats|assert" | head -30
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -5
for cfg in "LS_KNN_SMALL_TILED=0" "LS_KNN_SMALL_TILED=1"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 > gpurun_out/r2/ab6_$tag.json 2> gpurun_out/r2/ab6_$tag.err
  tail -3 gpurun_out/r2/ab6_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab6_$tag.json"))
    st=d["stages_ms"]
    print("$cfg", round(d["value"]), round(d["ms_per_step"],3), {k:v for k,v in st.items() if "global" in k or "edgeconv[5" in k or "edgeconv[6" in k})
except Exception as e:
    print("$cfg FAILED", e)
PY
done
