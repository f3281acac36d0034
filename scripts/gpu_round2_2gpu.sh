set -x
O=gpurun_out/r2final
mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_configs.py -q -k "c4_sharded" 2>&1 | tail -3 | tee $O/pytest_c4_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu_r02.json 2> $O/bench_2gpu_r02.err; tail -c 400 $O/bench_2gpu_r02.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_2gpu_r02.json 2> $O/bench_ref_2gpu_r02.err; tail -c 300 $O/bench_ref_2gpu_r02.err
python - <<PY
import json
d=json.load(open("$O/bench_2gpu_r02.json"))
print(d["value"], d["ms_per_step"], d["checked"], d["checks"], json.dumps(d["strong_c4"])[:900])
PY
