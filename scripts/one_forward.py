"""A few encoder forwards of the bench batch (profiling target for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import livingscenes_b200 as ls
dev = torch.device("cuda:0")
sd, _ = bench.load_state_dict()
model = ls.Shape_Prior.from_state_dict(sd).to(dev).eval()
x, _ = bench.make_scene_batch(bench.PAIRS_PER_GPU, 101)
x = x.to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n):
    model.encode_packed(x)
torch.cuda.synchronize()
