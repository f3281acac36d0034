mkdir -p gpurun_out/r2
for md in 0 4; do
env LS_KNN_TC_MIN_D=$md timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab18_$md.json 2> gpurun_out/r2/ab18_$md.err
tail -2 gpurun_out/r2/ab18_$md.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab18_$md.json"))
    st=d["stages_ms"]
    print("min_d $md", round(d["value"]), round(d["ms_per_step"],3), d["checked"], {k:v for k,v in st.items() if "[0]" in k})
except Exception as e:
    print("FAILED", e)
PY
done
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
