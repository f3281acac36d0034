# usage: gpu_round2_ngpu.sh N  (run under gpurun --gpus N): weak-scaling bench line (incl. the strong-scaling config[3] section)
set -x
N=$1
O=gpurun_out/r2final
mkdir -p $O
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${N}gpu_r02.json 2> $O/bench_${N}gpu_r02.err; tail -c 300 $O/bench_${N}gpu_r02.err
python - <<PY
import json
d=json.load(open("$O/bench_${N}gpu_r02.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], d["checked"], {k:v for k,v in d["strong_c4"].items() if k in ("ms_per_step","instances_per_s","gathered_table_bit_identical_to_1gpu_encode")})
PY
