"""Latency of the ragged batched FPS front end (encode_fps) on a 3RScan-like batch."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from livingscenes_b200.ops import farthest_point_sample_masked, farthest_point_sample
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, Nmax = 32, 40000
pc = torch.randn(B, 3, Nmax, generator=g).to(dev)
n_valid = torch.randint(2000, Nmax, (B,), generator=g)
mask = (torch.arange(Nmax)[None, :] < n_valid[:, None]).to(dev)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
t1 = timeit(lambda: farthest_point_sample_masked(pc, mask, 1024))
def loop():
    for b in range(B):
        farthest_point_sample(pc[b:b + 1, :, :int(n_valid[b])].contiguous(), 1024)
t2 = timeit(loop, 2)
print(f"B={B} Nmax={Nmax} (valid 2000..40000): ls_fps_masked one launch {t1:.2f} ms; per-instance loop {t2:.2f} ms")
