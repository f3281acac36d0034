"""One pass of the SDF decoder (8 instances x 16384 query points = 131072 columns): profiling target for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import livingscenes_b200 as ls
from livingscenes_b200 import synthetic as S
dev = torch.device("cuda:0")
sd, _ = bench.load_state_dict()
model = ls.Shape_Prior.from_state_dict(sd).to(dev).eval()
x = S.synth_instances(8, 1024, 77).to(dev)
codes = model.encode(x)
q = S.sdf_queries(codes["s"].cpu(), codes["t"].cpu(), 16384, 1239).to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n):
    out = model.decoder(q, None, codes, return_sdf=True)
torch.cuda.synchronize()
print(float(out.abs().mean()))
