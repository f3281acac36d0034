import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import livingscenes_b200 as ls
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
z0 = torch.randn(128, 256, generator=g).to(dev); z1 = torch.randn(128, 256, generator=g).to(dev)
sizes = [32] * 4
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
print("sequential_matcher_batched 4 x (32+32): %.1f us" % (1e3 * timeit(lambda: ls.sequential_matcher_batched(z0, z1, sizes, sizes))))
gr = torch.cuda.CUDAGraph()
ls.sequential_matcher_batched(z0, z1, sizes, sizes); torch.cuda.synchronize()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): ls.sequential_matcher_batched(z0, z1, sizes, sizes)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
with torch.cuda.graph(gr): out = ls.sequential_matcher_batched(z0, z1, sizes, sizes)
print("graph replay: %.1f us" % (1e3 * timeit(lambda: gr.replay())))
