"""Does spatially coherent point order speed up the gather phase?  Time the encoder stages on the bench batch
with the input points in random order vs sorted by the Morton code of xyz (layers 0-1 keep the input order)."""
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

def morton(x):
    mn, mx = x.min(-1, keepdim=True)[0], x.max(-1, keepdim=True)[0]
    u = ((x - mn) / (mx - mn + 1e-9) * 1023).long().clamp(0, 1023)  # [B,3,N]
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(u[:, 0]) | (spread(u[:, 1]) << 1) | (spread(u[:, 2]) << 2)
    perm = code.argsort(-1)
    return torch.gather(x, 2, perm[:, None, :].expand(-1, 3, -1))

def run(xin, tag):
    xin = xin.to(dev)
    acc = {}
    _lib.profile_enable(True)
    for k in range(6):
        model.encoder.run(xin, normalize=True)
        torch.cuda.synchronize()
        if k >= 2:
            for name, layer, ms in _lib.profile_read():
                acc.setdefault((name, layer), []).append(ms)
    _lib.profile_enable(False)
    out = {f"{n}[{l}]": round(statistics.mean(v), 4) for (n, l), v in sorted(acc.items())}
    print(json.dumps({"tag": tag, "total": round(sum(out.values()), 3), **{k: v for k, v in out.items() if "knn" in k}}))

run(x, "random order")
run(morton(x), "morton order")
