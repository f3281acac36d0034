mkdir -p gpurun_out/r2
for pf in 0 2 4 8 16; do
env LS_EDGE_PREFETCH=$pf timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab13_pf$pf.json 2> gpurun_out/r2/ab13_pf$pf.err
tail -2 gpurun_out/r2/ab13_pf$pf.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab13_pf$pf.json"))
    st=d["stages_ms"]
    print("prefetch $pf", round(d["value"]), round(d["ms_per_step"],3), d["checked"], {k:v for k,v in st.items() if "edgeconv" in k})
except Exception as e:
    print("FAILED", e)
PY
done
