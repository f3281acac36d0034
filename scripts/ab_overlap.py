"""A/B: encoder forward of the bench batch with and without side-stream overlap (CUDA events, eager + graph)."""
import os, sys, statistics
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
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)
for on in (False, True):
    _lib.set_overlap(on)
    eager = timeit(lambda: model.encode_packed(x))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        model.encode_packed(x)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        out = model.encode_packed(x)
    graph = timeit(lambda: g.replay())
    print(f"overlap={on}: eager {eager:.3f} ms, graph {graph:.3f} ms")
