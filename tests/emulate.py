"""Test-only torch emulation of the ALGEBRA the CUDA kernels implement (folded point-level GEMMs +
gather, sqrt-free VN activation, attention-pool normalisation shortcuts), driven by the very arrays
``VecDGCNN_att._fold`` / ``DeepSDF_Decoder._pack`` hand to the C ABI.  It lets the CPU test-suite
check the host-side weight folding against the oracle without a GPU.  Not product code."""
from __future__ import annotations

import math

import torch


def _vn_act(q, k, oms):
    """q, k [...,3] -> q - (1-slope) * min(<q,k>,0) / max(|k|^2, 1e-24) * k   (ls_common.cuh: vn_act)."""
    n2 = (k * k).sum(-1, keepdim=True)
    dt = (q * k).sum(-1, keepdim=True)
    t = oms * dt.clamp(max=0.0) / n2.clamp_min(1e-24)
    return q - t * k


def _table(W, f):
    """W [R,C], f [B,C,3,N] -> point-major table [B,N,R,3]."""
    return torch.einsum("rc,bcan->bnra", W, f)


def encoder_emulated(enc, x, knn_idx, fps_idx):
    """enc: livingscenes_b200.VecDGCNN_att (CPU params are fine), x [B,3,N]; graph teacher-forced."""
    F = {k: v.float() for k, v in enc._fold().items()}
    oms = 1.0 - enc.leak_neg_slope
    B = x.shape[0]
    src_f = x.unsqueeze(1)
    n_fps = 0
    feats = []
    for i in range(enc.num_layers):
        Co = enc.feat_dim[i]
        if i in enc.down_sample_layers:
            sel = fps_idx[n_fps].long()
            n_fps += 1
            dst_f = torch.gather(src_f, 3, sel[:, None, None, :].expand(B, src_f.shape[1], 3, sel.shape[1]))
        else:
            dst_f = src_f
        idx = knn_idx[i].long()  # [B,Nd,16]
        Nd = idx.shape[1]
        if i == 0:
            W = F["l0.w0"]  # [2][Co][3]
            xyz = src_f[:, 0].transpose(1, 2)  # [B,N,3]
            nn = torch.gather(xyz[:, None].expand(B, Nd, xyz.shape[1], 3), 2, idx[..., None].expand(B, Nd, 16, 3))
            xd = xyz[:, :, None, :].expand_as(nn)
            h = xyz / xyz.norm(dim=-1, keepdim=True).clamp_min(1e-12)
            cr = torch.linalg.cross(h[:, :, None, :].expand_as(nn), nn, dim=-1)
            df = nn - xd
            q = torch.einsum("oc,bnkca->bnkoa", W[0], torch.stack([cr, df, xd], 3))
            k = torch.einsum("oc,bnkca->bnkoa", W[1], torch.stack([cr, df, xd], 3))
            out = _vn_act(q, k, oms).mean(2)  # [B,Nd,Co,3]
        else:
            att = i >= enc.atten_start_layer
            Ps = _table(F[f"l{i}.w_src"], src_f)  # [B,Ns,R,3]
            Pd = _table(F[f"l{i}.w_dst"], dst_f)  # [B,Nd,R',3]
            G = torch.gather(Ps[:, None].expand(B, Nd, *Ps.shape[1:]), 2,
                             idx[..., None, None].expand(B, Nd, 16, Ps.shape[2], 3))  # [B,Nd,16,R,3]
            blk = lambda t, p: t[..., p * Co:(p + 1) * Co, :]
            vq = blk(G, 0) + blk(Pd, 0)[:, :, None]
            vk = blk(G, 1) + blk(Pd, 1)[:, :, None]
            vv = _vn_act(vq, vk, oms)  # [B,Nd,16,Co,3]
            if not att:
                out = vv.mean(2)
            else:
                kq = blk(G, 2) + blk(Pd, 2)[:, :, None]
                kk = blk(G, 3) + blk(Pd, 3)[:, :, None]
                ko = _vn_act(kq, kk, oms)
                qv = _vn_act(blk(Pd, 4), blk(Pd, 5), oms)  # [B,Nd,Co,3]
                ell = qv.norm(dim=-1, keepdim=True)
                Lq = (ell ** 2).sum(2, keepdim=True).sqrt().clamp_min(1e-12)
                qq = qv / ell.clamp_min(1e-12) * (ell / Lq)
                Lk = (ko ** 2).sum((-1, -2)).sqrt().clamp_min(1e-12)  # [B,Nd,16]
                r = (ko * qq[:, :, None]).sum(-1)  # [B,Nd,16,Co]
                r = r.reshape(B, Nd, 16, Co // 16, 16).sum(-1)  # heads
                logit = r / Lk[..., None] / math.sqrt(48.0)
                a = torch.softmax(logit, dim=2)  # over the 16 edges
                a = a[..., None].expand(B, Nd, 16, Co // 16, 16).reshape(B, Nd, 16, Co)
                out = (a[..., None] * vv).sum(2)
        out = out.permute(0, 2, 3, 1).contiguous()  # [B,Co,3,Nd]
        if enc.use_res_global_conv and i >= enc.res_global_start_layer:
            g = out.mean(-1)  # [B,Co,3]
            raw = torch.einsum("rc,bcan->bran", F[f"l{i}.w_g1"], out) + \
                torch.einsum("rc,bca->bra", F[f"l{i}.w_g2"], g)[..., None]
            out = _vn_act(raw[:, :Co].permute(0, 1, 3, 2), raw[:, Co:].permute(0, 1, 3, 2), oms).permute(0, 1, 3, 2)
        feats.append(out)
        src_f = out
    C = enc.c_dim
    raw = torch.einsum("rc,bcan->bran", F["w_conv_c"], src_f)  # [B,C+1,3,Nl]
    xh = _vn_act(raw[:, :C].permute(0, 1, 3, 2), raw[:, C:].permute(0, 1, 3, 2).expand(B, C, -1, 3), oms).mean(2)
    ell = xh.norm(dim=-1, keepdim=True)
    L = (ell ** 2).sum(1, keepdim=True).sqrt().clamp_min(1e-12)
    z_so3 = xh / ell.clamp_min(1e-12) * (ell / L)
    scale = ell.squeeze(-1).mean(1) * enc.scale_factor
    u = torch.einsum("kc,bka->bca", F["w_inv_t"], xh)
    ul = u.norm(dim=-1, keepdim=True)
    UL = (ul ** 2).sum(1, keepdim=True).sqrt().clamp_min(1e-12)
    z_inv = ((u / ul.clamp_min(1e-12) * (ul / UL)) * z_so3).sum(-1)
    f0 = torch.einsum("kc,bka->bca", F["w_fc0_t"], xh)
    h = C // 2
    net = _vn_act(f0[:, :h], f0[:, h:], oms)
    v = torch.einsum("c,bca->ba", F["w_lin1"], net) + torch.einsum("c,bca->ba", F["w_short"], xh)
    w2 = float(enc.fc_center.act2.lin_dir.weight.detach().reshape(-1)[0])
    center = _vn_act(v, w2 * v, oms)
    if enc.center_pred_scale:
        center = center * enc.scale_factor
    return center.unsqueeze(1), scale, z_so3, z_inv, feats


# ----------------------------------------------------------------------------------------------------------
# CPU emulation of the tensor-core kNN path (csrc/ls_knn_tc.cu + k_knn_rerank): ranking value from a 3xTF32
# product, threshold from group minima, candidate set, hybrid re-rank (exact distances for ambiguous runs only).
def _tf32_rn(x):
    import numpy as np
    u = x.view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _tf32_trunc(x):
    import numpy as np
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tc_knn_emulated(q, s, K=16, group_cols=256, kappa_scale=1.0):
    """q [D,Nq], s [D,Ns] float32 numpy -> (idx [Nq,K] int64, stats).  Mirrors the kernels' decisions; the exact
    distance is the sequential fp32 FMA form of the brute-force kernel."""
    import numpy as np

    D, Nq = q.shape
    Ns = s.shape[1]
    ns = np.zeros(Ns, np.float32)
    nq = np.zeros(Nq, np.float32)
    for d in range(D):
        ns = (s[d] * s[d] + ns).astype(np.float32)
        nq = (q[d] * q[d] + nq).astype(np.float32)
    sh, qh = _tf32_rn(s), _tf32_rn(q)
    sl, ql = _tf32_trunc((s - sh).astype(np.float32)), _tf32_trunc((q - qh).astype(np.float32))
    f64 = np.float64
    dot = (qh.T.astype(f64) @ sh.astype(f64) + qh.T.astype(f64) @ sl.astype(f64) + ql.T.astype(f64) @ sh.astype(f64))
    dt = (ns[None, :] - 2 * dot.astype(np.float32)).astype(np.float32)
    kappa = kappa_scale * (2 * D * 2.0 ** -23 + 2.0 ** -16)
    e2 = (2 * kappa * (nq + ns.max())).astype(np.float32)

    def exact(qi, si):  # sequential fp32 FMA chain (float64 product of float32 factors rounds like an FMA)
        acc = np.float32(0)
        for d in range(D):
            df = np.float32(q[d, qi] - s[d, si])
            acc = np.float32(f64(df) * f64(df) + f64(acc))
        return acc

    idx = np.zeros((Nq, K), np.int64)
    n_cand, n_exact = [], []
    for i in range(Nq):
        m = np.full(32, np.inf, np.float32)
        stash = []
        for g0 in range(0, Ns, group_cols):
            cols = np.arange(g0, min(g0 + group_cols, Ns))
            np.minimum.at(m, cols % 32, dt[i, cols])
            thr = np.sort(m)[15] + e2[i]
            stash = [c for c in stash if dt[i, c] <= thr] + [c for c in cols if dt[i, c] <= thr]
        cand = sorted(stash, key=lambda c: (dt[i, c], c))
        n_cand.append(len(cand))
        # runs of candidates chained by gaps <= 2E that reach into the first K ranks get exact distances
        run_start, need = [], []
        for r, c in enumerate(cand):
            adj_prev = r > 0 and np.float32(dt[i, c] - dt[i, cand[r - 1]]) <= e2[i]
            run_start.append(run_start[-1] if adj_prev else r)
        for r, c in enumerate(cand):
            adj_next = r + 1 < len(cand) and np.float32(dt[i, cand[r + 1]] - dt[i, c]) <= e2[i]
            adj_prev = run_start[r] != r
            need.append((adj_prev or adj_next) and run_start[r] <= K - 1)
        n_exact.append(sum(need))
        order = sorted(range(len(cand)), key=lambda r: (run_start[r], exact(i, cand[r]) if need[r] else np.float32(0), cand[r]))
        idx[i] = [cand[r] for r in order[:K]]
    return idx, {"candidates_mean": float(np.mean(n_cand)), "candidates_max": int(max(n_cand)),
                 "exact_per_query_mean": float(np.mean(n_exact))}
