"""Seeded synthetic inputs and stand-in weights (no reference counterpart): the workload generators of bench.py,
scripts/ and the tests.  Pure torch on the CPU; nothing here computes any part of the hot path.

  synth_instances    SURVEY.md 8d synthetic clouds (noisy ellipsoidal shells; the round-1 fixtures use these)
  synth_parts        asymmetric composite objects (random boxes of one "furniture-like" assembly): every instance has
                     a unique pose and a distinctive shape, so a planted permutation / SE(3) is recoverable
  random_rotations   uniform random proper rotations
  random_state_dict  seeded weights with the shipped checkpoint's keys and shapes
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

# encoder hyper-parameters of the shipped checkpoint (weights/files_backup/model_config.yaml:142-171)
SHIPPED_ENCODER_CFG = dict(
    c_dim=256, num_layers=7, feat_dim=[32, 32, 64, 64, 128, 256, 512],
    down_sample_layers=[2, 4, 5], down_sample_factor=[2, 4, 4],
    atten_start_layer=2, atten_multi_head_c=16, use_res_global_conv=True,
    res_global_start_layer=2, num_knn=16, scale_factor=64000.0, leak_neg_slope=0.2,
    use_dg=True, center_pred=True, center_pred_scale=True,
)


# --------------------------------------------------------------------------- weights
def random_state_dict(seed: int = 0, cfg: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """Seeded stand-in weights with the shipped checkpoint's keys and shapes
    (kaiming_uniform(a=sqrt 5) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)), vec_layers.py:117)."""
    cfg = dict(SHIPPED_ENCODER_CFG, **(cfg or {}))
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def uni(name, out_c, in_c, fan_in=None):
        bnd = 1.0 / math.sqrt(fan_in or in_c)
        sd[name] = (torch.rand(out_c, in_c, generator=g) * 2 - 1) * bnd

    def lna(prefix, cin, cout, shared=False):
        uni(prefix + ".lin.weight", cout, cin)
        uni(prefix + ".act.lin_dir.weight", 1 if shared else cout, cout)

    fd = cfg["feat_dim"]
    for i in range(cfg["num_layers"]):
        cin = 3 if i == 0 else 2 * fd[i - 1]
        lna(f"encoder.V_list.{i}", cin, fd[i])
        if i >= cfg["atten_start_layer"]:
            lna(f"encoder.K_list.{i}", cin, fd[i])
            lna(f"encoder.Q_list.{i}", fd[i - 1], fd[i])
        if cfg["use_res_global_conv"] and i >= cfg["res_global_start_layer"]:
            lna(f"encoder.global_conv_list.{i - cfg['res_global_start_layer']}", 2 * fd[i], fd[i])
    c = cfg["c_dim"]
    lna("encoder.conv_c", fd[-1], c, shared=True)
    uni("encoder.fc_inv.weight", c, c)
    lna("encoder.fc_center.fc0", c, c // 2)
    uni("encoder.fc_center.lin1.weight", 1, c // 2)
    uni("encoder.fc_center.act2.lin_dir.weight", 1, 1)
    uni("encoder.fc_center.shortcut.weight", 1, c)
    dims = [513, 768, 768, 768, 255, 768, 768, 768, 768, 1]
    ins = [513, 768, 768, 768, 768, 768, 768, 768, 768]
    for l in range(9):
        o, i_ = dims[l + 1], ins[l]
        bnd = 1.0 / math.sqrt(i_)
        v = (torch.rand(o, i_, generator=g) * 2 - 1) * bnd
        bias = (torch.rand(o, generator=g) * 2 - 1) * bnd
        if l < 8:
            sd[f"decoder.lin{l}.bias"] = bias
            # a non-trivial gain so that weight-norm folding is actually exercised
            sd[f"decoder.lin{l}.weight_g"] = v.norm(dim=1, keepdim=True) * (0.75 + 0.5 * torch.rand(o, 1, generator=g))
            sd[f"decoder.lin{l}.weight_v"] = v
        else:
            sd[f"decoder.lin{l}.weight"] = v
            sd[f"decoder.lin{l}.bias"] = bias
    return sd


# --------------------------------------------------------------------------- synthetic inputs
def synth_instances(B: int, N: int, seed: int) -> torch.Tensor:
    """SURVEY.md section 8d synthetic clouds: noisy ellipsoidal shells, [B,3,N] fp32.
    Directions uniform on S^2, radii 0.3 + 0.2 U(0,1), per-axis scale (1.0, 0.6, 0.8) times a
    per-instance factor U(0.5, 1.5), N(0, 0.005^2) noise, translation U(-2, 2)^3."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(B, N, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-9)
    r = 0.3 + 0.2 * torch.rand(B, N, 1, generator=g)
    ax = torch.tensor([1.0, 0.6, 0.8])[None, None, :]
    sc = 0.5 + torch.rand(B, 1, 1, generator=g)
    # a few low-frequency bumps so instances are distinguishable by shape
    bump = 1.0 + 0.25 * torch.sin(3.0 * d[..., :1] + 6.28 * torch.rand(B, 1, 1, generator=g)) \
        * torch.cos(2.0 * d[..., 1:2] + 6.28 * torch.rand(B, 1, 1, generator=g))
    p = d * r * ax * sc * bump + 0.005 * torch.randn(B, N, 3, generator=g)
    p = p + (torch.rand(B, 1, 3, generator=g) * 4 - 2)
    return p.transpose(1, 2).contiguous().float()


def synth_parts(B: int, N: int, seed: int, noise: float = 0.002) -> torch.Tensor:
    """Asymmetric composite objects, [B,3,N] fp32.  An instance is the surface of 4-7 axis-aligned boxes of random
    size hung on a random "seat" slab (legs of different length/offset, a back, an arm on ONE side), scaled by
    U(0.5, 1.5) and translated by U(-2, 2)^3.  Unlike the ellipsoidal shells of ``synth_instances`` the shapes have no
    rotational symmetry, so the equivariant code pins the rotation and the invariant code separates the instances
    (bench.py / scripts/run_configs.py plant a permutation + SE(3) and check that the path recovers them)."""
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(B, N, 3)
    for b in range(B):
        n_box = int(torch.randint(4, 8, (1,), generator=g))
        u = lambda *s: torch.rand(*s, generator=g)
        seat = torch.tensor([0.35, 0.30, 0.04]) * (0.7 + 0.6 * u(3))
        half = [seat]
        cen = [torch.zeros(3)]
        for k in range(n_box - 1):
            kind = k % 3
            if kind == 0:    # a leg below the seat, its own length and footprint position
                h = torch.tensor([0.03, 0.03, 0.10 + 0.25 * float(u(1))]) * (0.7 + 0.6 * u(3))
                c = torch.stack([(u(1)[0] * 2 - 1) * seat[0] * 0.9, (u(1)[0] * 2 - 1) * seat[1] * 0.9, -seat[2] - h[2]])
            elif kind == 1:  # a back / panel above one edge
                h = torch.tensor([0.30, 0.03, 0.15 + 0.25 * float(u(1))]) * (0.6 + 0.8 * u(3))
                c = torch.stack([(u(1)[0] * 2 - 1) * 0.1, seat[1] * (0.6 + 0.4 * u(1)[0]), seat[2] + h[2]])
            else:            # an arm / shelf on the +x side only
                h = torch.tensor([0.04, 0.20, 0.05]) * (0.6 + 0.8 * u(3))
                c = torch.stack([seat[0] * (0.7 + 0.3 * u(1)[0]), (u(1)[0] * 2 - 1) * 0.1, seat[2] + 0.05 + 0.2 * u(1)[0]])
            half.append(h)
            cen.append(c)
        half, cen = torch.stack(half), torch.stack(cen)
        # area-weighted surface sampling: pick a box, then one of its 6 faces
        fa = torch.stack([half[:, 1] * half[:, 2], half[:, 0] * half[:, 2], half[:, 0] * half[:, 1]], 1)  # face areas /4
        w = fa.repeat(1, 2).reshape(-1)
        pick = torch.multinomial(w / w.sum(), N, replacement=True, generator=g)
        bi, fi = pick // 6, pick % 6
        p = (u(N, 3) * 2 - 1) * half[bi]
        ax, sgn = fi % 3, (fi // 3).float() * 2 - 1
        p[torch.arange(N), ax] = sgn * half[bi, ax]
        p = (p + cen[bi]) * (0.5 + float(u(1)))
        out[b] = p + noise * torch.randn(N, 3, generator=g) + (u(1, 3) * 4 - 2)
    return out.transpose(1, 2).contiguous().float()


def random_rotations(B: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(B, 3, 3, generator=g, dtype=torch.float64)
    Q, R = torch.linalg.qr(A)
    Q = Q * torch.sign(torch.diagonal(R, dim1=1, dim2=2))[:, None, :]
    Q[:, :, 0] *= torch.det(Q)[:, None]
    return Q.float()


def lcg_uniform(n: int, seed: int) -> torch.Tensor:
    """n doubles in [0,1) from a 64-bit LCG evaluated in closed form per index with numpy uint64 arithmetic
    (Knuth MMIX multiplier, splitmix-style finaliser): bit-reproducible on every platform / torch version, so
    large query sets (config C5: 4 x 100 000 points) are regenerated from the seed instead of stored."""
    import numpy as np

    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = i * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return torch.from_numpy((z >> np.uint64(11)).astype(np.float64) / float(1 << 53))


def sdf_queries(code_s: torch.Tensor, code_t: torch.Tensor, M: int, seed: int) -> torch.Tensor:
    """Config C5 queries: M points per instance, uniform in the 1.1-padded unit cube of the instance's canonical
    frame (mesh_extractor2.py:100) mapped to the world by q*s + t.  code_s [B], code_t [B,1,3] -> [B,M,3] fp32."""
    B = code_s.shape[0]
    u = lcg_uniform(B * M * 3, seed).reshape(B, M, 3).float()
    return ((u - 0.5) * 1.1) * code_s.reshape(B, 1, 1).cpu() + code_t.reshape(B, 1, 3).cpu()
