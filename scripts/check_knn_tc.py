"""A/B of the kNN graph: tensor-core filter + exact re-rank vs exact brute force (same library, same inputs)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import livingscenes_b200 as ls
from livingscenes_b200 import _lib

dev = torch.device("cuda:0")
sd, _ = bench.load_state_dict()
model = ls.Shape_Prior.from_state_dict(sd).to(dev).eval()
enc = model.encoder
for (B, N) in ((2, 512), (8, 1024), (3, 2048), (5, 1000), (256, 1024)):
    from livingscenes_b200 import synthetic as R
    x = R.synth_instances(B, N, 99 + B).to(dev)
    outs = {}
    for name, on, ks in (("exact", False, 1.0), ("tc", True, 1.0), ("tc_overflow", True, 1e6)):
        _lib.set_knn_tensor_cores(on, ks)
        r = enc.run(x, normalize=True, taps=True)
        torch.cuda.synchronize()
        outs[name] = r
    _lib.set_knn_tensor_cores(True, 1.0)
    for name in ("tc", "tc_overflow"):
        bad = []
        for i, (a, b_) in enumerate(zip(outs["exact"]["knn_idx"], outs[name]["knn_idx"])):
            nd = int((a != b_).any(-1).sum())
            bad.append(nd)
        dz = float((outs["exact"]["z_inv"] - outs[name]["z_inv"]).abs().max())
        print(f"B={B} N={N} {name}: differing kNN rows per layer {bad}, max |dz_inv| {dz:.2e}", flush=True)
print("done")
