# usage: gpu_ab_variants.sh name1 name2 ...   (variants/NAME.so built by scripts/build_variant.sh)
mkdir -p gpurun_out/r2
cp livingscenes_b200/_ls_b200.so /tmp/_orig.so
for v in "$@"; do
  cp variants/$v.so livingscenes_b200/_ls_b200.so
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "knn" 2>&1 | tail -1
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > gpurun_out/r2/abv_$v.json 2> gpurun_out/r2/abv_$v.err
  tail -2 gpurun_out/r2/abv_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/abv_$v.json"))
    st=d["stages_ms"]
    grp={}
    for k,x in st.items():
        g=k.split("[")[0]; grp[g]=round(grp.get(g,0)+x,3)
    print("VARIANT $v", round(d["value"]), round(d["ms_per_step"],3), grp)
    print("   ", {k:v for k,v in st.items() if "filter" in k or "rerank" in k})
except Exception as e:
    print("VARIANT $v FAILED", e)
PY
done
cp /tmp/_orig.so livingscenes_b200/_ls_b200.so
