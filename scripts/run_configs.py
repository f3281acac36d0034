"""Run the BASELINE.json configs 2-5 on one GPU and print one JSON object (profiles/r01/configs.json).
config 2: B=256 x N=1024 encoder forward            (also bench.py's workload)
config 3: 2 x 32 instances N=2048: encode + sequential match + Kabsch, SE(3) error vs the planted transform
config 5: 64 instances x 100k SDF queries
Timing: CUDA events, 3 warm-ups, median of 10."""
import json, os, sys, statistics, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import livingscenes_b200 as ls
from livingscenes_b200 import _lib
from livingscenes_b200 import synthetic as R

dev = torch.device("cuda:0")
SHIPPED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "livingscenes_b200", "_weights", "shipped_fp32.pt")
sd = torch.load(SHIPPED, map_location="cpu", weights_only=True) if os.path.exists(SHIPPED) else R.random_state_dict(0)
model = ls.Shape_Prior.from_state_dict(sd).to(dev).eval()
solver = ls.More_Solver(model)


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


out = {"weights": "shipped" if os.path.exists(SHIPPED) else "random"}
# ---- config 2
x = R.synth_parts(256, 1024, 1235).to(dev)
ms = timed(lambda: model.encode(x))
out["config2_encode_B256_N1024"] = {"ms": ms, "instances_per_s": 256 / ms * 1e3}
# ---- config 3
g = torch.Generator().manual_seed(1236)
xa = R.synth_parts(32, 2048, 1236)          # asymmetric objects; the rescan is an exact rigid, permuted copy
perm = torch.randperm(32, generator=g)
Rg = R.random_rotations(32, 1237)
tg = torch.randn(32, 3, 1, generator=g)
xb = Rg @ xa[perm] + tg
xa_d, xb_d = xa.to(dev), xb.to(dev)
ms = timed(lambda: solver.solve_scene_pair(xa_d, xb_d))
res = solver.solve_scene_pair(xa_d, xb_d)
inv = torch.argsort(perm)
m0 = res["matches"]["matches0"].cpu()
ok = (m0 == inv)
R_gt = Rg[inv]  # transform of ref instance i onto its rescan copy
rre = ls.rotation_error(res["R"].cpu(), R_gt).reshape(-1)
# the planted translation acts on world coordinates: x_b = R x_a + t
rte = ls.translation_error(res["t"].cpu(), tg[inv])
out["config3_pair_2x32_N2048"] = {"ms": ms, "instances_per_s": 64 / ms * 1e3, "match_recall": float(ok.float().mean()),
                                   "rre_deg_median": float(rre[ok].median()) if ok.any() else None,
                                   "rte_median": float(rte[ok].median()) if ok.any() else None,
                                   "rre_deg_max": float(rre[ok].max()) if ok.any() else None}
assert out["config3_pair_2x32_N2048"]["match_recall"] == 1.0 and out["config3_pair_2x32_N2048"]["rre_deg_median"] < 1.0, \
    out["config3_pair_2x32_N2048"]  # the planted permutation and poses must be recovered
# ---- config 5
codes = model.encode(x[:64])
M = 100_000
q = ((torch.rand(64, M, 3, generator=torch.Generator().manual_seed(5)) - 0.5) * 1.1).to(dev) * codes["s"][:, None, None] + codes["t"]
ms = timed(lambda: model.decoder(q, None, codes, return_sdf=True), n=5, warm=2)
flops = 64 * M * 2 * (257 * 768 + 2 * 768 * 768 + 768 * 255 + 512 * 768 + 3 * 768 * 768 + 768)
out["config5_sdf_64x100k"] = {"ms": ms, "points_per_s": 64 * M / ms * 1e3, "TFLOPs_fp32_equiv": flops / ms * 1e-9,
                              "tensor_TFLOPs_tf32_issued": 3 * flops / ms * 1e-9, "gemm": "tcgen05 TS form (variant 3)"}
_lib.set_gemm_variant(1)
ms1 = timed(lambda: model.decoder(q, None, codes, return_sdf=True), n=5, warm=2)
_lib.set_gemm_variant(3)
out["config5_sdf_64x100k_gemm_variant1"] = {"ms": ms1, "points_per_s": 64 * M / ms1 * 1e3, "tensor_TFLOPs_tf32_issued": 3 * flops / ms1 * 1e-9}
_lib.set_tensor_cores(False)
ms2 = timed(lambda: model.decoder(q[:8], None, {k: v[:8] for k, v in codes.items()}, return_sdf=True), n=3, warm=1)
_lib.set_tensor_cores(True)
out["config5_sdf_simt_fp32_8x100k"] = {"ms": ms2, "points_per_s": 8 * M / ms2 * 1e3}
print(json.dumps(out))
