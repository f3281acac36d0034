set -x
mkdir -p gpurun_out/r2
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_configs.py -q -k "c4_sharded" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/bench_2gpu.json 2> gpurun_out/r2/bench_2gpu.err; tail -c 600 gpurun_out/r2/bench_2gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2/bench_2gpu.json"))
print(d["value"], d["ms_per_step"], d["checked"], d["checks"], json.dumps(d["strong_c4"])[:900])
PY
