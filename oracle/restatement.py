"""CPU restatement of the LivingScenes per-instance inference hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- the checker for the CUDA path
and the timed CPU "port" baseline; never imported by livingscenes_b200.

Written from the arithmetic specification in SURVEY.md Appendix A, in the
reference's own (edge-level, eager fp32) computational form, as plain functions
over a flat weight dict (keys ``encoder.*`` / ``decoder.*`` exactly as the
shipped checkpoint names them after stripping ``network_dict.``).  Each function
cites the reference lines it follows (paths relative to /root/reference).

Checked against the reference's own modules by oracle/make_golden.py (in the
build container) and against the committed fixtures by tests/test_oracle.py.
The pytorch3d boundary is "parity unpinned" (oracle/p3d_shim.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import p3d_shim

# Shipped encoder hyper-parameters: weights/files_backup/model_config.yaml:142-171
SHIPPED_ENCODER_CFG = dict(
    c_dim=256, num_layers=7, feat_dim=[32, 32, 64, 64, 128, 256, 512],
    down_sample_layers=[2, 4, 5], down_sample_factor=[2, 4, 4],
    atten_start_layer=2, atten_multi_head_c=16, use_res_global_conv=True,
    res_global_start_layer=2, num_knn=16, scale_factor=64000.0, leak_neg_slope=0.2,
    use_dg=True, center_pred=True, center_pred_scale=True,
)
# Shipped decoder: model_config.yaml:105-140 (latent 256 + pe 257 = 513 wide input)
SHIPPED_DECODER_CFG = dict(latent_size=256, pe_dim=257, dims=[768] * 8, latent_in=[4])

EPS_NRM = 1e-12  # F.normalize default eps


# --------------------------------------------------------------------------- VN primitives
def _nrm(v: torch.Tensor, dim: int) -> torch.Tensor:
    """v / max(|v|_2, 1e-12) along ``dim``  (F.normalize, used at vec_layers.py:27,29,263)."""
    return v / v.norm(dim=dim, keepdim=True).clamp_min(EPS_NRM)


def vn_linear(W: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Channel mixing of 3-vectors, no bias: out[b,o,a,...] = sum_c W[o,c] v[b,c,a,...]
    (VecLinear.forward so3 vector-only path, vec_layers.py:121-134)."""
    return torch.einsum("oc,bc...->bo...", W, v)


def vn_act(W_dir: torch.Tensor, q: torch.Tensor, slope: float) -> torch.Tensor:
    """VN leaky-ReLU with a learned direction (VecActivation.forward so3, vec_layers.py:241-268).
    W_dir is [C,C], or [1,C] when the direction is shared across channels."""
    k = vn_linear(W_dir, q)
    kh = _nrm(k, 2)
    p = (q * kh).sum(dim=2, keepdim=True)
    return q - p * kh + torch.nn.functional.leaky_relu(p, slope) * kh


def vn_lna(Wd: Dict[str, torch.Tensor], prefix: str, v: torch.Tensor, slope: float) -> torch.Tensor:
    """VecLNA = VecLinear then VecActivation (vec_layers.py:523-534)."""
    q = vn_linear(Wd[prefix + ".lin.weight"], v)
    return vn_act(Wd[prefix + ".act.lin_dir.weight"], q, slope)


def cevn(x: torch.Tensor) -> torch.Tensor:
    """channel_equi_vec_normalize (vec_layers.py:24-31): unit directions times the
    channel-normalised lengths."""
    ln = x.norm(dim=2, keepdim=True)
    return _nrm(x, 2) * _nrm(ln, 1)


# --------------------------------------------------------------------------- graph ops
def knn_graph(dst_f: torch.Tensor, src_f: torch.Tensor, K: int) -> torch.Tensor:
    """Feature-space kNN of every dst point among the src points (vec_dgcnn_atten.py:136-141).
    dst_f [B,C,3,Nd], src_f [B,C,3,Ns] -> idx [B,Nd,K] int64 (ascending distance, ties -> lower idx).
    The distance runs over d = c*3 + axis, the order reshape(B, C*3, N) gives (:138)."""
    B, C, _, Nd = dst_f.shape
    Ns = src_f.shape[-1]
    q = dst_f.reshape(B, C * 3, Nd).transpose(1, 2)
    s = src_f.reshape(B, C * 3, Ns).transpose(1, 2)
    _, idx, _ = p3d_shim.knn_points(q, s, K=K, return_nn=False)
    return idx


def fps_indices(xyz: torch.Tensor, n_out: int) -> torch.Tensor:
    """Farthest point sampling on xyz [B,3,N] -> idx [B,n_out] (vec_dgcnn_atten.py:169)."""
    _, idx = p3d_shim.sample_farthest_points(xyz.transpose(1, 2), K=n_out)
    return idx


def _gather_points(f: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """f [B,C,3,N], idx [B,M] -> [B,C,3,M]  (vec_dgcnn_atten.py:173)."""
    B, C, _, _ = f.shape
    return torch.gather(f, 3, idx[:, None, None, :].expand(B, C, 3, idx.shape[1]))


def _edge_feature(src_f, dst_f, idx, first_layer: bool) -> torch.Tensor:
    """Edge feature cat[(cross), nn - dst, dst] -> [B,2C(or 3),3,Nd,K] (vec_dgcnn_atten.py:142-160).
    The cross term uses the vector axis (dim=2); the reference's dim-less torch.cross
    picks dim 0 when B == 3 -- treated as a reference bug, B == 3 is excluded from parity."""
    B, C, _, Ns = src_f.shape
    Nd, K = idx.shape[1], idx.shape[2]
    nn = torch.gather(src_f[:, :, :, None, :].expand(B, C, 3, Nd, Ns), 4,
                      idx[:, None, None, :, :].expand(B, C, 3, Nd, K))
    dst = dst_f[..., None].expand_as(nn)
    if first_layer:
        xdir = _nrm(src_f, 2)[..., None].expand_as(nn)
        cr = torch.linalg.cross(xdir, nn, dim=2)
        return torch.cat([cr, nn - dst, dst], 1)
    return torch.cat([nn - dst, dst], 1)


# --------------------------------------------------------------------------- encoder
def encoder_forward(Wd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: Optional[dict] = None,
                    trace: Optional[dict] = None, force: Optional[dict] = None):
    """VecDGCNN_att.forward (vec_dgcnn_atten.py:177-252).  x [B,3,N] fp32 ->
    (center [B,1,3], scale [B], z_so3 [B,c_dim,3], z_inv [B,c_dim]).

    ``trace`` (a dict) collects per-layer ``knn_idx``, ``fps_idx``, ``feat`` (layer
    outputs) and ``src_f``/``dst_f`` (layer inputs) for kernel-level parity tests.
    ``force`` may carry ``knn_idx`` / ``fps_idx`` lists to teacher-force the graph.
    """
    cfg = dict(SHIPPED_ENCODER_CFG, **(cfg or {}))
    P = "encoder."
    slope = cfg["leak_neg_slope"]
    K = cfg["num_knn"]
    hc = cfg["atten_multi_head_c"]
    B, _, N = x.shape
    assert x.dtype == torch.float32
    src_xyz = x.unsqueeze(1)
    src_f = x.unsqueeze(1)
    if trace is not None:
        trace.update(knn_idx=[], fps_idx=[], feat=[], src_f=[], dst_f=[])
    n_fps = 0
    for i in range(cfg["num_layers"]):
        if i in cfg["down_sample_layers"]:  # :188-191, :163-175
            factor = cfg["down_sample_factor"][cfg["down_sample_layers"].index(i)]
            n_new = src_xyz.shape[-1] // factor
            if force is not None and "fps_idx" in force:
                sel = force["fps_idx"][n_fps]
            else:
                sel = fps_indices(src_xyz.squeeze(1), n_new)
            n_fps += 1
            if trace is not None:
                trace["fps_idx"].append(sel)
            dst_xyz = _gather_points(src_xyz, sel)
            dst_f = _gather_points(src_f, sel)
        else:
            dst_xyz, dst_f = src_xyz, src_f
        if force is not None and "knn_idx" in force:
            idx = force["knn_idx"][i]
        else:
            idx = knn_graph(dst_f, src_f, K)
        if trace is not None:
            trace["knn_idx"].append(idx)
            trace["src_f"].append(src_f)
            trace["dst_f"].append(dst_f)
        y = _edge_feature(src_f, dst_f, idx, first_layer=(i == 0))
        if i < cfg["atten_start_layer"]:  # :202-204
            out = vn_lna(Wd, f"{P}V_list.{i}", y, slope).mean(-1)
        else:  # :205-219
            kk = cevn(vn_lna(Wd, f"{P}K_list.{i}", y, slope))
            qq = cevn(vn_lna(Wd, f"{P}Q_list.{i}", dst_f, slope))
            vv = vn_lna(Wd, f"{P}V_list.{i}", y, slope)
            qk = (kk * qq[..., None]).sum(2)  # B,C,Nd,K
            Bq, C, Nd, Kk = qk.shape
            logit = qk.reshape(Bq, C // hc, hc, Nd, Kk).sum(2, keepdim=True) / math.sqrt(3 * hc)
            att = torch.softmax(logit, dim=-1).expand(-1, -1, hc, -1, -1).reshape(Bq, C, 1, Nd, Kk)
            out = (att * vv).sum(-1)
        if cfg["use_res_global_conv"] and i >= cfg["res_global_start_layer"]:  # :222-225
            g = out.mean(-1, keepdim=True).expand_as(out)
            j = i - cfg["res_global_start_layer"]
            out = vn_lna(Wd, f"{P}global_conv_list.{j}", torch.cat([out, g], 1), slope)
        if trace is not None:
            trace["feat"].append(out)
        src_xyz, src_f = dst_xyz, out
    # head :231-252
    xh = vn_lna(Wd, f"{P}conv_c", src_f, slope).mean(-1)  # B,c_dim,3 (shared direction)
    z_so3 = cevn(xh)
    scale = xh.norm(dim=-1).mean(1) * cfg["scale_factor"]
    z_dual = vn_linear(Wd[f"{P}fc_inv.weight"], xh[..., None]).squeeze(-1)
    z_inv = (cevn(z_dual) * z_so3).sum(-1)
    if not cfg["center_pred"]:
        return scale, z_so3, z_inv
    # VecResBlock (vec_layers.py:631-672), so3 mode, vector-only, last_activate=True
    v = xh[..., None]
    net = vn_lna(Wd, f"{P}fc_center.fc0", v, slope)
    dv = vn_linear(Wd[f"{P}fc_center.lin1.weight"], net)
    vs = vn_linear(Wd[f"{P}fc_center.shortcut.weight"], v)
    center = vn_act(Wd[f"{P}fc_center.act2.lin_dir.weight"], vs + dv, slope).squeeze(-1)
    if cfg["center_pred_scale"]:
        center = center * cfg["scale_factor"]
    return center, scale, z_so3, z_inv


def scale0(xc: torch.Tensor) -> torch.Tensor:
    """scale_0 = mean of the 5 largest entries of the flattened N x N distance matrix of
    the centred cloud xc [B,3,N] (model_utils.py:175-176; torch.cdist default mode)."""
    B = xc.shape[0]
    pts = xc.transpose(-1, -2)
    d = torch.cdist(pts, pts)
    return d.reshape(B, -1).topk(5, dim=-1)[0].mean(-1)


def encode(Wd, x: torch.Tensor, cfg: Optional[dict] = None, trace: Optional[dict] = None) -> dict:
    """Shape_Prior.encode (model_utils.py:165-197), use_double=False."""
    x = x.float()
    mu = x.mean(-1)
    xc = x - mu[..., None]
    s0 = scale0(xc)
    xn = xc / s0[:, None, None]
    if trace is not None:
        trace["scale0"] = s0
        trace["x_norm"] = xn
    center, scale, z_so3, z_inv = encoder_forward(Wd, xn, cfg, trace)
    return {"z_so3": z_so3, "z_inv": z_inv, "s": s0 * scale, "t": (center.squeeze(1) + mu).unsqueeze(1)}


# --------------------------------------------------------------------------- matching
def sequential_match(m0: torch.Tensor, m1: torch.Tensor) -> dict:
    """Greedy matcher (lib_more/matcher_new.py:109-139): rescale by max + 1e-5, take the
    first (row-major) entry equal to the max, retire its row and column; min(n, m) rounds."""
    a = m0 / m0.norm(dim=1, keepdim=True).clamp_min(EPS_NRM)
    b = m1 / m1.norm(dim=1, keepdim=True).clamp_min(EPS_NRM)
    n, m = a.shape[0], b.shape[0]
    rows, cols = list(range(n)), list(range(m))
    out0 = torch.full((n,), -1, dtype=torch.int64, device=m0.device)
    out1 = torch.full((m,), -1, dtype=torch.int64, device=m0.device)
    S = a @ b.T
    for _ in range(min(n, m)):
        S = S / (S.max() + 1e-5)
        flat = int((S == S.max()).reshape(-1).nonzero()[0, 0])
        r, c = divmod(flat, S.shape[1])
        out0[rows[r]] = cols[c]
        out1[cols[c]] = rows[r]
        keep_r = [i for i in range(S.shape[0]) if i != r]
        keep_c = [j for j in range(S.shape[1]) if j != c]
        S = S[keep_r][:, keep_c]
        del rows[r], cols[c]
    return {"matches0": out0, "matches1": out1}


def mutual_nn_match(desc0: torch.Tensor, desc1: torch.Tensor) -> dict:
    """nn_matcher (matcher_new.py:85-105): desc [1,D,n] / [1,D,m], cosine top-1 both ways,
    keep mutual pairs, -1 otherwise; outputs squeezed."""
    a = desc0 / desc0.norm(dim=1, keepdim=True).clamp_min(EPS_NRM)
    b = desc1 / desc1.norm(dim=1, keepdim=True).clamp_min(EPS_NRM)
    sim = torch.einsum("bdn,bdm->bnm", a, b)
    f0 = sim.argmax(dim=2)
    f1 = sim.argmax(dim=1)
    n, m = f0.shape[1], f1.shape[1]
    ok0 = torch.gather(f1, 1, f0) == torch.arange(n, device=f0.device)[None]
    m0 = torch.where(ok0, f0, torch.full_like(f0, -1))
    back = torch.gather(m0, 1, f1)
    ok1 = back == torch.arange(m, device=f0.device)[None]
    m1 = torch.where(ok1, f1, torch.full_like(f1, -1))
    return {"matches0": m0.squeeze(), "matches1": m1.squeeze()}


# --------------------------------------------------------------------------- pose
def kabsch(x1: torch.Tensor, x2: torch.Tensor, weights: Optional[torch.Tensor] = None,
           normalize_w: bool = True, eps: float = 1e-7):
    """Weighted Kabsch (lib_more/pose_estimation.py:29-121): x1, x2 [b,n,3] ->
    R [b,3,3], t [b,3,1], residual norms [b,n].  Orientation fix R = V diag(1,1,det(V U^T)) U^T."""
    b, n, _ = x1.shape
    w = torch.ones(b, n, dtype=x1.dtype, device=x1.device) if weights is None else weights
    if normalize_w:
        w = w / (w.sum(dim=1, keepdim=True) + eps)
    w = w.unsqueeze(2)
    den = w.sum(dim=1, keepdim=True) + eps
    mu1 = (w * x1).sum(dim=1, keepdim=True) / den
    mu2 = (w * x2).sum(dim=1, keepdim=True) / den
    c1, c2 = x1 - mu1, x2 - mu2
    cov = c1.transpose(1, 2) @ (w * c2)
    U, _, Vh = torch.linalg.svd(cov)
    V = Vh.transpose(1, 2)
    det = torch.det(V @ U.transpose(1, 2))
    Dm = torch.diag_embed(torch.stack([torch.ones_like(det), torch.ones_like(det), det], 1))
    R = V @ Dm @ U.transpose(1, 2)
    t = mu2.transpose(1, 2) - R @ mu1.transpose(1, 2)
    res = ((R @ x1.transpose(1, 2) + t).transpose(1, 2) - x2).norm(dim=2)
    return R, t, res


# --------------------------------------------------------------------------- decoder
def decoder_weights(Wd: Dict[str, torch.Tensor]) -> List[tuple]:
    """Effective (W, b) of the 9 linears; weight_norm(dim=0): W = g * v / |v|_row
    (deepsdf_decoder.py:51-58, torch.nn.utils.weight_norm)."""
    out = []
    for l in range(9):
        p = f"decoder.lin{l}."
        if p + "weight_v" in Wd:
            v, g = Wd[p + "weight_v"], Wd[p + "weight_g"]
            W = v * (g / v.norm(dim=1, keepdim=True))
        else:
            W = Wd[p + "weight"]
        out.append((W, Wd[p + "bias"]))
    return out


def sdf_decode(Wd, query: torch.Tensor, code: dict, latent_in=(4,)) -> torch.Tensor:
    """FieldWrapper.forward, inner_deepsdf branch (model_utils.py:230-263) followed by
    DeepSDF_Decoder.forward(phase='val') (deepsdf_decoder.py:78-123).  query [B,M,3] -> sdf [B,M]."""
    B, M, _ = query.shape
    q = (query - code["t"]) / code["s"][:, None, None]
    inner = torch.einsum("bmd,bcd->bmc", q, code["z_so3"])
    length = q.norm(dim=-1, keepdim=True)
    u = torch.cat([code["z_inv"][:, None, :].expand(-1, M, -1), inner, length], -1).reshape(B * M, -1)
    h = u
    Ws = decoder_weights(Wd)
    for l, (W, bias) in enumerate(Ws):
        if l in latent_in:
            h = torch.cat([h, u], 1)
        h = h @ W.T + bias
        if l < len(Ws) - 1:
            h = torch.relu(h)
    return torch.tanh(h).reshape(B, M)


# --------------------------------------------------------------------------- synthetic inputs / weights
# The seeded generators live in the package (livingscenes_b200/synthetic.py: pure-torch CPU helpers shared by
# bench.py, scripts/ and the tests) so that nothing outside tests / smoke / the CPU-baseline leg imports oracle/.


# --------------------------------------------------------------------------- secondary matchers
def sinkhorn_match(desc0: torch.Tensor, desc1: torch.Tensor, iters: int = 100, alpha: float = 1.0,
                   match_threshold: float = 0.0) -> dict:
    """sinkhorn_matcher (matcher_new.py:11-71): desc [1,D,n] / [1,D,m]; cosine scores / sqrt(D), log-space optimal
    transport with a dustbin row and column (alpha), mutual arg-max on the [n,m] block, exp(score) > threshold."""
    a = desc0 / desc0.norm(dim=1, keepdim=True).clamp_min(EPS_NRM)
    b = desc1 / desc1.norm(dim=1, keepdim=True).clamp_min(EPS_NRM)
    S = torch.einsum("bdn,bdm->bnm", a, b)[0] / math.sqrt(desc0.shape[1])
    n, m = S.shape
    Z = torch.full((n + 1, m + 1), float(alpha), dtype=S.dtype, device=S.device)
    Z[:n, :m] = S
    norm = -math.log(n + m)
    log_mu = torch.full((n + 1,), norm, dtype=S.dtype, device=S.device)
    log_nu = torch.full((m + 1,), norm, dtype=S.dtype, device=S.device)
    log_mu[n] = math.log(m) + norm
    log_nu[m] = math.log(n) + norm
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v[None, :], dim=1)
        v = log_nu - torch.logsumexp(Z + u[:, None], dim=0)
    Z = Z + u[:, None] + v[None, :] - norm
    blk = Z[:n, :m]
    v0, i0 = blk.max(dim=1)
    i1 = blk.max(dim=0).indices
    ar_n, ar_m = torch.arange(n, device=S.device), torch.arange(m, device=S.device)
    mutual0 = i1[i0] == ar_n
    mutual1 = i0[i1] == ar_m
    valid0 = mutual0 & (torch.where(mutual0, v0.exp(), torch.zeros_like(v0)) > match_threshold)
    valid1 = mutual1 & valid0[i1]
    return {"matches0": torch.where(valid0, i0, torch.full_like(i0, -1)),
            "matches1": torch.where(valid1, i1, torch.full_like(i1, -1))}


def residual_seq_match(src_codes: dict, tgt_codes: dict, use_sim: bool) -> dict:
    """sim3_seq_matcher (use_sim, matcher_new.py:142-185) / eq_seq_matcher (:188-230): the greedy rounds of
    sequential_match on  cos / (res + 1e-5)  or  1 / (res + 1e-5),  res[i,j] = mean Kabsch residual of the
    equivariant codes z_so3 src[i] -> tgt[j]."""
    a = src_codes["z_inv"] / src_codes["z_inv"].norm(dim=1, keepdim=True).clamp_min(EPS_NRM)
    b = tgt_codes["z_inv"] / tgt_codes["z_inv"].norm(dim=1, keepdim=True).clamp_min(EPS_NRM)
    n, m = a.shape[0], b.shape[0]
    res = torch.zeros(n, m, dtype=a.dtype, device=a.device)
    for i in range(n):
        _, _, r = kabsch(src_codes["z_so3"][i][None].repeat_interleave(m, dim=0), tgt_codes["z_so3"])
        res[i] = r.mean(dim=1)
    S = (a @ b.T) / (res + 1e-5) if use_sim else 1.0 / (res + 1e-5)
    rows, cols = list(range(n)), list(range(m))
    out0 = torch.full((n,), -1, dtype=torch.int64, device=a.device)
    out1 = torch.full((m,), -1, dtype=torch.int64, device=a.device)
    for _ in range(min(n, m)):
        S = S / (S.max() + 1e-5)
        flat = int((S == S.max()).reshape(-1).nonzero()[0, 0])
        r, c = divmod(flat, S.shape[1])
        out0[rows[r]] = cols[c]
        out1[cols[c]] = rows[r]
        keep_r = [i for i in range(S.shape[0]) if i != r]
        keep_c = [j for j in range(S.shape[1]) if j != c]
        S = S[keep_r][:, keep_c]
        del rows[r], cols[c]
    return {"matches0": out0, "matches1": out1}


from livingscenes_b200.synthetic import random_rotations, random_state_dict, synth_instances  # noqa: E402,F401
