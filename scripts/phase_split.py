"""Split k_knn_edge time per layer into phase 1 (kNN) and phase 2 (EdgeConv+pool): run the bench batch
free-running and with the graph teacher-forced (phase 2 only), per-stage CUDA events (ls_profile_*)."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import livingscenes_b200 as ls
from livingscenes_b200 import _lib

dev = torch.device("cuda:0")
sd, _ = bench.load_state_dict()
model = ls.Shape_Prior.from_state_dict(sd).to(dev).eval()
x, _ = bench.make_scene_batch(bench.PAIRS_PER_GPU, 101)
x = x.to(dev)
enc = model.encoder
r = enc.run(x, normalize=True, taps=True)
idx = [t.clone() for t in r["knn_idx"]]
torch.cuda.synchronize()

def timed(force):
    acc = {}
    _lib.profile_enable(True)
    for k in range(6):
        enc.run(x, normalize=True, force_knn_idx=idx if force else None)
        torch.cuda.synchronize()
        if k >= 2:
            for name, layer, ms in _lib.profile_read():
                acc.setdefault((name, layer), []).append(ms)
    _lib.profile_enable(False)
    return {f"{n}[{l}]": round(statistics.mean(v), 4) for (n, l), v in sorted(acc.items())}

free, forced = timed(False), timed(True)
out = {}
for k in free:
    if k.startswith("knn_edgeconv"):
        out[k] = {"total": free[k], "phase2_only": forced.get(k), "phase1_est": round(free[k] - forced.get(k, 0), 4)}
print(json.dumps({"split": out, "free": free}, indent=1))
