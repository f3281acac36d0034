set -x
mkdir -p gpurun_out/r2
for cfg in "LS_KNN_SMALL_TILED=0" "LS_KNN_SMALL_TILED=1"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/ab6_$tag.json 2> gpurun_out/r2/ab6_$tag.err
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
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench_c.json 2> gpurun_out/r2/bench_c.err; tail -c 800 gpurun_out/r2/bench_c.err; head -c 300 gpurun_out/r2/bench_c.json
timeout 300 python bench.py --workload sdf > gpurun_out/r2/bench_sdf.json 2> gpurun_out/r2/bench_sdf.err; tail -c 500 gpurun_out/r2/bench_sdf.err; cat gpurun_out/r2/bench_sdf.json | cut -c1-1500
