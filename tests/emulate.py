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
