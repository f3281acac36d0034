set -x
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm_persistent or vn_linear" 2>&1 | tail -15
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "=== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench_a.json 2> gpurun_out/r2/bench_a.err; tail -c 600 gpurun_out/r2/bench_a.err
echo "=== A/B"
for cfg in "LS_WAVE_MB=0 LS_GEMM_VARIANT=1" "LS_WAVE_MB=0 LS_GEMM_VARIANT=2" "LS_WAVE_MB=28 LS_GEMM_VARIANT=1" "LS_WAVE_MB=14 LS_GEMM_VARIANT=2" "LS_WAVE_MB=56 LS_GEMM_VARIANT=2"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 > gpurun_out/r2/ab_$tag.json 2> gpurun_out/r2/ab_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/ab_$tag.json"))
    st=d["stages_ms"]
    agg={}
    for k,v in st.items():
        agg[k.split("[")[0]]=agg.get(k.split("[")[0],0)+v
    print("$cfg", round(d["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in agg.items()})
except Exception as e:
    print("$cfg FAILED", e)
PY
done
