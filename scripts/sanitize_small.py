"""Small invocation of every new kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import livingscenes_b200 as ls
from livingscenes_b200 import synthetic as S
from livingscenes_b200.ops import (farthest_point_sample, farthest_point_sample_masked, iterative_closest_point,
                                   knn_graph_cm_tc)
dev = torch.device("cuda:0")
model = ls.Shape_Prior.from_state_dict(S.random_state_dict(0)).to(dev).eval()
x = S.synth_instances(2, 600, 3).to(dev)
out = model.encode(x)
q = torch.randn(1, 24, 300, device=dev)
idx, d2, nc = knn_graph_cm_tc(q, torch.randn(1, 24, 700, device=dev))
pc = torch.randn(2, 3, 9000, device=dev)
farthest_point_sample(pc, 64)
mask = torch.rand(2, 9000, device=dev) < 0.5
farthest_point_sample_masked(pc, mask, 64)
iterative_closest_point(torch.randn(2, 300, 3, device=dev), torch.randn(2, 400, 3, device=dev), max_iterations=5)
# round 2: the SDF decoder forward / backward (k_gemm_tc3 with 256-row tiles, ReLU-mask epilogue), a batch large enough for the
# persistent GEMMs (k_gemm_tc2 point-major and channel-major epilogues, k_gemm_tc3 tables), the matchers and Kabsch
codes = {k: v for k, v in out.items()}
qq = (torch.rand(2, 700, 3, device=dev) - 0.5).requires_grad_(True)
sdf = model.decoder(qq, None, codes, return_sdf=True)
sdf.sum().backward()
xb = S.synth_instances(40, 1024, 5).to(dev)
ob = model.encode(xb)
m = ls.sequential_matcher_batched(ob["z_inv"][:20].contiguous(), ob["z_inv"][20:].contiguous(), [20], [20])
torch.cuda.synchronize()
print("ok", float(out["z_inv"].abs().sum()), int(nc.max()), float(sdf.abs().sum()), float(ob["z_inv"].abs().sum()), int(m["matches0"].sum()))
