"""Per-layer knn_filter / knn_edgeconv stage times of the bench batch (quick A/B of kernel variants)."""
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
acc = {}
_lib.profile_enable(True)
for k in range(6):
    model.encoder.run(x, normalize=True)
    torch.cuda.synchronize()
    if k >= 2:
        for name, layer, ms in _lib.profile_read():
            acc.setdefault((name, layer), []).append(ms)
_lib.profile_enable(False)
out = {f"{n}[{l}]": round(statistics.mean(v), 4) for (n, l), v in sorted(acc.items())}
tot = sum(out.values())
print(json.dumps({"tag": os.environ.get("LS_NVCC_EXTRA", ""), "total": round(tot, 3), **{k: v for k, v in out.items() if "knn" in k}}))
