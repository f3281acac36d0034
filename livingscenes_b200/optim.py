"""Inference-time optimisation loops of lib_more/more_solver.py on the differentiable CUDA decoder.

  optimize_code        more_solver.py:191-228  Adam on (z_inv, t, z_so3) against MSE(sdf(observed points), 0), 200 steps
  refine_registration  more_solver.py:118-179  Adam on an SE(3) element against SmoothL1(sdf(g . src)) + Sinkhorn(g . src, tgt)

The SDF and its gradients come from ``ls_sdf_decode`` / ``ls_sdf_backward`` through ``FieldWrapper`` (decoder.py).  Adam,
the LR schedules and the losses are torch's own (the reference uses exactly these classes).  Two third-party pieces of
the reference are NOT in this container (pytorch3d aside): ``torchlie`` (the SE(3) parameter) and ``geomloss``
(``SamplesLoss('sinkhorn', p=2)``).  They are restated below from their published algorithms -- parity UNPINNED for
these two: no reference-side vectors exist and neither library can be run here.
"""
from __future__ import annotations

import math

import torch

from .pose_estimation import kabsch_transformation_estimation


# --------------------------------------------------------------------------------------- _optimize_code
def optimize_code(model, code, pc, n_steps: int = 200):
    """more_solver.py:191-228.  ``pc`` [B,N,3] (already FPS'ed), ``code`` dict -> optimised code dict.

    Quirk reproduced: the reference keeps ``best_code = {k: code[k].detach()}``; ``detach()`` shares storage with the
    parameters Adam keeps updating in place, so what it returns is the code after the LAST step, whatever the loss
    history was (as long as one step had loss < 100).  ``s`` is not optimised."""
    code = {k: v.detach().clone() for k, v in code.items()}
    groups = [{"params": [code["z_inv"]], "lr": 1e-5}, {"params": [code["t"]], "lr": 1e-4},
              {"params": [code["z_so3"]], "lr": 5e-4}]
    for g in groups:
        g["params"][0].requires_grad_(True)
    opt = torch.optim.Adam(groups)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[160], gamma=0.1)
    loss_fn = torch.nn.MSELoss()
    losses = []
    for _ in range(n_steps):
        opt.zero_grad()
        sdf = model.decoder(pc, None, code, return_sdf=True)
        loss = loss_fn(sdf, torch.zeros_like(sdf))
        loss.backward()
        opt.step()
        sched.step()
        losses.append(loss.detach())
    out = {k: v.detach() for k, v in code.items()}
    out["loss_history"] = torch.stack(losses)
    return out


# --------------------------------------------------------------------------------------- SE(3) helpers
def _hat(w):
    z = torch.zeros_like(w[..., 0])
    return torch.stack([torch.stack([z, -w[..., 2], w[..., 1]], -1), torch.stack([w[..., 2], z, -w[..., 0]], -1),
                        torch.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def so3_exp(w):
    """Rodrigues formula, [b,3] -> [b,3,3] (series near 0)."""
    th = w.norm(dim=-1, keepdim=True).unsqueeze(-1)
    K = _hat(w)
    small = th < 1e-4
    th_s = torch.where(small, torch.ones_like(th), th)
    a = torch.where(small, 1 - th ** 2 / 6, torch.sin(th_s) / th_s)
    b = torch.where(small, 0.5 - th ** 2 / 24, (1 - torch.cos(th_s)) / th_s ** 2)
    return torch.eye(3, device=w.device, dtype=w.dtype) + a * K + b * (K @ K)


def rotation_geodesic(R1, R2):
    """roma.rotmat_geodesic_distance: the rotation angle of R1^T R2 in RADIANS."""
    tr = torch.diagonal(R1.transpose(-1, -2) @ R2, dim1=-2, dim2=-1).sum(-1)
    return torch.acos(((tr - 1) * 0.5).clamp(-1, 1))


# --------------------------------------------------------------------------------------- Sinkhorn divergence
def _softmin(eps, C, h):
    """-eps * logsumexp_j (h_j - C_ij / eps):  C [b,n,m], h [b,m] -> [b,n]."""
    return -eps * torch.logsumexp(h[:, None, :] - C / eps, dim=2)


def sinkhorn_divergence(x, y, blur: float = 0.05, scaling: float = 0.5):
    """geomloss.SamplesLoss('sinkhorn', p=2, blur=0.05, scaling=0.5, debias=True) for uniform weights, restated from
    the published algorithm (Feydy et al. 2019: symmetric log-domain Sinkhorn with epsilon-scaling from diameter^2
    down to blur^2, debiased S_eps = OT(a,b) - OT(a,a)/2 - OT(b,b)/2, gradients through the last extrapolation step
    only).  x [b,n,3], y [b,m,3] -> [b].  UNPINNED: geomloss is not installable here."""
    b, n, _ = x.shape
    m = y.shape[1]
    cost = lambda u, v: 0.5 * torch.cdist(u, v) ** 2
    la = torch.full((b, n), -math.log(n), device=x.device, dtype=x.dtype)
    lb = torch.full((b, m), -math.log(m), device=x.device, dtype=x.dtype)
    with torch.no_grad():
        xd, yd = x.detach(), y.detach()
        mins = torch.minimum(xd.amin(1), yd.amin(1))
        maxs = torch.maximum(xd.amax(1), yd.amax(1))
        diameter = float((maxs - mins).norm(dim=-1).max().clamp_min(blur))
        eps_list = [diameter ** 2] + [math.exp(e) for e in
                                      torch.arange(2 * math.log(diameter), 2 * math.log(blur), 2 * math.log(scaling)).tolist()] + [blur ** 2]
        Cxy, Cyx, Cxx, Cyy = cost(xd, yd), cost(yd, xd), cost(xd, xd), cost(yd, yd)
        eps = eps_list[0]
        f_ba, g_ab = _softmin(eps, Cxy, lb), _softmin(eps, Cyx, la)
        f_aa, g_bb = _softmin(eps, Cxx, la), _softmin(eps, Cyy, lb)
        for eps in eps_list:
            ft_ba = _softmin(eps, Cxy, lb + g_ab / eps)
            gt_ab = _softmin(eps, Cyx, la + f_ba / eps)
            ft_aa = _softmin(eps, Cxx, la + f_aa / eps)
            gt_bb = _softmin(eps, Cyy, lb + g_bb / eps)
            f_ba, g_ab = 0.5 * (f_ba + ft_ba), 0.5 * (g_ab + gt_ab)
            f_aa, g_bb = 0.5 * (f_aa + ft_aa), 0.5 * (g_bb + gt_bb)
    # last extrapolation with autograd enabled on the sample positions
    eps = eps_list[-1]
    F_ba = _softmin(eps, cost(x, y.detach()), lb + g_ab / eps)
    G_ab = _softmin(eps, cost(y, x.detach()), la + f_ba / eps)
    F_aa = _softmin(eps, cost(x, x.detach()), la + f_aa / eps)
    G_bb = _softmin(eps, cost(y, y.detach()), lb + g_bb / eps)
    return ((F_ba - F_aa) * la.exp()).sum(1) + ((G_ab - G_bb) * lb.exp()).sum(1)


# --------------------------------------------------------------------------------------- optim=True registration
def refine_registration(model, pc1, pc2, code1, code2, R, t, cfg):
    """more_solver.py:118-179 for a batch of independent pairs: pc1, pc2 [B,N,3] (FPS'ed), codes of both, initial
    R [B,3,3], t [B,3,1] (pc1 -> pc2).  Per pair: pick the direction whose shared code explains its own points worse
    (:120-138), then ``n_steps`` Adam steps (lr ``step_size.so3``, MultiStepLR [300,340,380] x0.1) on the SE(3) element
    against SmoothL1(sdf(g . src), 0) + Sinkhorn(g . src, tgt); the best-loss element is returned, inverted where the
    direction was flipped.  The SE(3) parameter is updated on the manifold, g <- exp(-step) g with Adam moments kept on
    the 6-vector left-tangent gradient (torchlie's LieTensor is not available here: restated, unpinned).  The early
    stop compares ``roma.rotmat_geodesic_distance`` (radians) with ``early_stop_threshold`` (10), as the reference does."""
    B = pc1.shape[0]
    dev = pc1.device
    with torch.no_grad():
        e1 = model.decoder(pc1, None, code1, return_sdf=True).abs().mean(1)
        e2 = model.decoder(pc2, None, code2, return_sdf=True).abs().mean(1)
        fwd = e1 >= e2                                              # pc1 -> pc2 with code2 as the shared code
        Rb, tb, _, _ = kabsch_transformation_estimation(code2["z_so3"] + code2["t"], code1["z_so3"] + code1["t"])
        sel = lambda a, b: torch.where(fwd.reshape(-1, *([1] * (a.dim() - 1))), a, b)
        src, tgt = sel(pc1, pc2), sel(pc2, pc1)
        shared = {k: sel(code2[k], code1[k]).contiguous() for k in ("z_so3", "z_inv", "s", "t")}
        Rg, tg = sel(R, Rb).clone(), sel(t, tb).clone()
    R0 = Rg.clone()
    lr0 = float(cfg["step_size"]["so3"])
    n_steps, thr = int(cfg["n_steps"]), float(cfg["early_stop_threshold"])
    m = torch.zeros(B, 6, device=dev)
    v = torch.zeros(B, 6, device=dev)
    best_loss = torch.full((B,), 100.0, device=dev)
    best_R, best_t = Rg.clone(), tg.clone()
    smooth_l1 = torch.nn.SmoothL1Loss(reduction="none")
    b1, b2, eps_adam = 0.9, 0.999, 1e-8
    for i in range(n_steps):
        lr = lr0 * (0.1 ** sum(i >= ms for ms in (300, 340, 380)))
        Rv, tv = Rg.clone().requires_grad_(True), tg.clone().requires_grad_(True)
        q = src @ Rv.transpose(1, 2) + tv.transpose(1, 2)
        sdf = model.decoder(q, None, shared, return_sdf=True)
        loss = smooth_l1(sdf, torch.zeros_like(sdf)).mean(1) + sinkhorn_divergence(q, tgt)
        gR, gt = torch.autograd.grad(loss.sum(), (Rv, tv))
        with torch.no_grad():
            # left-tangent gradient of g -> exp(xi) g:  omega = vee(gR R^T - R gR^T) + t x gt,  nu = gt
            A = gR @ Rg.transpose(1, 2)
            om = torch.stack([A[:, 2, 1] - A[:, 1, 2], A[:, 0, 2] - A[:, 2, 0], A[:, 1, 0] - A[:, 0, 1]], -1)
            om = om + torch.linalg.cross(tg.squeeze(2), gt.squeeze(2), dim=-1)
            grad = torch.cat([om, gt.squeeze(2)], -1)
            m = b1 * m + (1 - b1) * grad
            v = b2 * v + (1 - b2) * grad * grad
            step = lr * (m / (1 - b1 ** (i + 1))) / ((v / (1 - b2 ** (i + 1))).sqrt() + eps_adam)
            dR = so3_exp(-step[:, :3])
            Rg = dR @ Rg
            tg = dR @ tg - step[:, 3:, None]
            better = loss.detach() < best_loss                      # the reference snapshots AFTER the step (:165-167)
            best_loss = torch.where(better, loss.detach(), best_loss)
            best_R = torch.where(better[:, None, None], Rg, best_R)
            best_t = torch.where(better[:, None, None], tg, best_t)
            if float(rotation_geodesic(Rg, R0).mean()) > thr:
                break
    with torch.no_grad():
        Ri = best_R.transpose(1, 2)
        R_out = torch.where(fwd[:, None, None], best_R, Ri)
        t_out = torch.where(fwd[:, None, None], best_t, -Ri @ best_t)
    return R_out, t_out
