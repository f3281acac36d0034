"""Weighted Kabsch SE(3) fit -- drop-in for lib_more/pose_estimation.py:29-121.

``kabsch_transformation_estimation(x1[b,n,3], x2[b,n,3], weights=None, normalize_w=True, eps=1e-7,
best_k=0, w_threshold=0) -> (R[b,3,3], t[b,3,1], res[b,n], flag)``; one warp per pair on the GPU
(weighted centroids and covariance by warp-shuffle reduction, 3x3 one-sided Jacobi SVD in registers,
proper-rotation fix, translation, residuals).  ``flag`` is the reference's "SVD failed" indicator;
the Jacobi solver cannot fail, so it is always False.  The RRE / RTE metrics stay plain torch.
"""
from __future__ import annotations

import torch

from . import _lib


@torch.no_grad()
def kabsch_transformation_estimation(x1, x2, weights=None, normalize_w=True, eps=1e-7, best_k=0, w_threshold=0):
    _lib.require_cuda(x1, "x1")
    dt = x1.dtype
    x1 = x1.detach().float()
    x2 = x2.detach().float()
    w = None if weights is None else weights.detach().float()
    if w is not None and normalize_w:
        # the reference normalises BEFORE best_k / w_threshold (pose_estimation.py:52-66)
        w = w / (w.sum(dim=1, keepdim=True) + eps)
        normalize_w = False
    if best_k > 0:
        if w is None:
            w = torch.full(x1.shape[:2], 1.0 / (x1.shape[1] + eps) if normalize_w else 1.0, device=x1.device)
            normalize_w = False
        ind = torch.topk(w[0], best_k, largest=True).indices  # reference: indices of batch 0 for all
        w, x1, x2 = w[:, ind], x1[:, ind], x2[:, ind]
    if w_threshold > 0:
        if w is None:  # the reference thresholds the normalised default weights 1 / (n + eps) too (pose_estimation.py:49-66)
            w = torch.full(x1.shape[:2], 1.0 / (x1.shape[1] + eps) if normalize_w else 1.0, device=x1.device)
            normalize_w = False
        w = torch.where(w < w_threshold, torch.zeros_like(w), w)
    x1, x2 = x1.contiguous(), x2.contiguous()
    b, n, _ = x1.shape
    dev = x1.device
    R = torch.empty(b, 3, 3, device=dev)
    t = torch.empty(b, 3, device=dev)
    res = torch.empty(b, n, device=dev)
    wc = None if w is None else w.contiguous()
    with torch.cuda.device(dev):
        rc = _lib.lib().ls_kabsch_batched(x1.data_ptr(), x2.data_ptr(), _lib.ptr(wc), b, n, int(bool(normalize_w)),
                                          float(eps), R.data_ptr(), t.data_ptr(), res.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, "ls_kabsch_batched")
        _lib.launch_count += 1
    return R.to(dt), t.unsqueeze(2).to(dt), res.to(dt), False


@torch.no_grad()
def kabsch_from_codes(code_a: dict, code_b: dict, matches0: torch.Tensor):
    """Pose of every matched instance pair straight from the embeddings, as
    More_Solver._solve_pairwise_registration does (more_solver.py:114-116):
    x1 = z_so3_a[i] + t_a[i], x2 = z_so3_b[m] + t_b[m], m = matches0[i] (identity where m < 0)."""
    za = code_a["z_so3"].detach().float().contiguous()
    zb = code_b["z_so3"].detach().float().contiguous()
    _lib.require_cuda(za, "code_a")
    n, c, _ = za.shape
    ta = code_a["t"].detach().float().reshape(n, 3).contiguous()
    tb = code_b["t"].detach().float().reshape(zb.shape[0], 3).contiguous()
    m = matches0.to(device=za.device, dtype=torch.int64).contiguous()
    R = torch.empty(n, 3, 3, device=za.device)
    t = torch.empty(n, 3, device=za.device)
    res = torch.empty(n, c, device=za.device)
    with torch.cuda.device(za.device):
        rc = _lib.lib().ls_kabsch_from_codes(za.data_ptr(), ta.data_ptr(), zb.data_ptr(), tb.data_ptr(), m.data_ptr(),
                                             n, c, R.data_ptr(), t.data_ptr(), res.data_ptr(), _lib.stream_ptr(za.device))
        _lib.check(rc, "ls_kabsch_from_codes")
        _lib.launch_count += 1
    return R, t.unsqueeze(2), res


def transformation_residuals(x1, x2, R, t):
    """pose_estimation.py:105-121 (plain torch; the CUDA Kabsch already returns these)."""
    return torch.norm((torch.matmul(R, x1.transpose(1, 2)) + t).transpose(1, 2) - x2, dim=2)


def rotation_error(R1, R2):
    """Geodesic angle between two batches of rotations, in degrees, shape [b,1]
    (pose_estimation.py:157-181): acos((trace(R1^T R2) - 1) / 2)."""
    tr = torch.diagonal(torch.matmul(R1.transpose(1, 2), R2), dim1=1, dim2=2).sum(-1)
    cos = ((tr - 1.0) * 0.5).clamp(-1.0, 1.0).unsqueeze(1)
    return torch.rad2deg(torch.acos(cos))


def translation_error(t1, t2):
    """Euclidean distance between translation vectors [b,3,1] (pose_estimation.py:183-196)."""
    return (t1 - t2).flatten(1).norm(dim=1)
