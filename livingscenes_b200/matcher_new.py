"""Instance matching on invariant codes -- drop-ins for lib_more/matcher_new.py.

  sequential_matcher(m0[n,256], m1[m,256])      matcher_new.py:109-139 (the matcher both evals select)
  nn_matcher(desc0[1,256,n], desc1[1,256,m])    matcher_new.py:85-105
Both return ``{'matches0', 'matches1'}`` int64 with -1 for unmatched, computed by one CTA per scene
pair on the GPU without any host synchronisation.  ``*_batched`` variants match many scene pairs in
one launch.
  sinkhorn_matcher(desc0[1,256,n], desc1[1,256,m])   matcher_new.py:45-71 (log-space OT, 100 iterations, dustbin)
  sim3_seq_matcher(src_codes, tgt_codes)             matcher_new.py:142-185 (cosine / mean Kabsch residual, greedy)
  eq_seq_matcher(src_codes, tgt_codes)               matcher_new.py:188-230 (1 / mean Kabsch residual, greedy)
The last three are selectable in ``More_Solver._solve_object_matching`` but not used by either eval.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib


def _offsets(sizes: Sequence[int]):
    arr = (C.c_int32 * (len(sizes) + 1))()
    acc = 0
    for i, s in enumerate(sizes):
        arr[i] = acc
        acc += int(s)
    arr[len(sizes)] = acc
    return arr, acc


@torch.no_grad()
def _match_batched(z0: torch.Tensor, z1: torch.Tensor, sizes0, sizes1, sequential: bool):
    _lib.require_cuda(z0, "z0")
    z0 = z0.detach().float().contiguous()
    z1 = z1.detach().float().contiguous()
    dim = z0.shape[1]
    off0, n0 = _offsets(sizes0)
    off1, n1 = _offsets(sizes1)
    assert n0 == z0.shape[0] and n1 == z1.shape[0] and z1.shape[1] == dim
    n_pairs = len(sizes0)
    dev = z0.device
    m0 = torch.empty(n0, dtype=torch.int64, device=dev)
    m1 = torch.empty(n1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        nbytes = C.c_size_t(0)
        _lib.check(_lib.lib().ls_match_workspace_bytes(off0, off1, n_pairs, C.byref(nbytes)), "ls_match_workspace_bytes")
        ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=dev)
        fn = _lib.lib().ls_seq_match if sequential else _lib.lib().ls_mutual_nn
        rc = fn(z0.data_ptr(), z1.data_ptr(), dim, off0, off1, n_pairs, m0.data_ptr(), m1.data_ptr(),
                ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "ls_seq_match" if sequential else "ls_mutual_nn")
        _lib.launch_count += 1
        ws.record_stream(torch.cuda.current_stream(dev))
    return m0, m1


def sequential_matcher(m0, m1):
    a, b = _match_batched(m0, m1, [m0.shape[0]], [m1.shape[0]], sequential=True)
    return {"matches0": a, "matches1": b}


def sequential_matcher_batched(z0, z1, sizes0, sizes1):
    """Many independent scene pairs in one launch: rows of z0/z1 are the concatenated instances,
    sizes0/sizes1 the per-pair instance counts.  Returned indices are local to each pair."""
    a, b = _match_batched(z0, z1, sizes0, sizes1, sequential=True)
    return {"matches0": a, "matches1": b}


def nn_matcher(desc0, desc1):
    assert desc0.dim() == 3 and desc0.shape[0] == 1, "nn_matcher takes [1,D,n] descriptors like the reference"
    a, b = _match_batched(desc0[0].T, desc1[0].T, [desc0.shape[2]], [desc1.shape[2]], sequential=False)
    return {"matches0": a.squeeze(), "matches1": b.squeeze()}


def nn_matcher_batched(z0, z1, sizes0, sizes1):
    a, b = _match_batched(z0, z1, sizes0, sizes1, sequential=False)
    return {"matches0": a, "matches1": b}


def _pair_workspace(n: int, m: int, dev):
    off0, off1 = (C.c_int32 * 2)(0, n), (C.c_int32 * 2)(0, m)
    nbytes = C.c_size_t(0)
    _lib.check(_lib.lib().ls_match_workspace_bytes(off0, off1, 1, C.byref(nbytes)), "ls_match_workspace_bytes")
    return torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=dev)


@torch.no_grad()
def sinkhorn_matcher(desc0, desc1, desc_dim=256, match_threshold=0.0, iters=100, alpha=1.0):
    """matcher_new.py:45-71.  ``desc_dim`` only scales the scores (1/sqrt(desc_dim)) in the reference; the kernel uses
    the descriptors' own dimension, which is what the reference passes (256)."""
    assert desc0.dim() == 3 and desc0.shape[0] == 1, "sinkhorn_matcher takes [1,D,n] descriptors like the reference"
    z0 = desc0[0].T.detach().float().contiguous()
    z1 = desc1[0].T.detach().float().contiguous()
    _lib.require_cuda(z0, "desc0")
    assert z0.shape[1] == desc_dim, "desc_dim must equal the descriptor dimension"
    n, m, dev = z0.shape[0], z1.shape[0], z0.device
    m0 = torch.empty(n, dtype=torch.int64, device=dev)
    m1 = torch.empty(m, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        ws = _pair_workspace(n, m, dev)
        rc = _lib.lib().ls_sinkhorn_match(z0.data_ptr(), z1.data_ptr(), z0.shape[1], n, m, int(iters), float(alpha),
                                          float(match_threshold), m0.data_ptr(), m1.data_ptr(), ws.data_ptr(), ws.numel(),
                                          _lib.stream_ptr(dev))
        _lib.check(rc, "ls_sinkhorn_match")
        _lib.launch_count += 1
        ws.record_stream(torch.cuda.current_stream(dev))
    return {"matches0": m0.squeeze(), "matches1": m1.squeeze()}


@torch.no_grad()
def _residual_seq_matcher(src_codes, tgt_codes, mode: int):
    from .pose_estimation import kabsch_transformation_estimation

    z0 = src_codes["z_inv"].detach().float().contiguous()
    z1 = tgt_codes["z_inv"].detach().float().contiguous()
    _lib.require_cuda(z0, "z_inv")
    n, m, dev = z0.shape[0], z1.shape[0], z0.device
    # res_mat[i, j] = mean residual of the Kabsch fit src z_so3[i] -> tgt z_so3[j]   (matcher_new.py:153-156)
    a = src_codes["z_so3"].detach().float()
    b = tgt_codes["z_so3"].detach().float()
    x1 = a[:, None].expand(n, m, *a.shape[1:]).reshape(n * m, *a.shape[1:]).contiguous()
    x2 = b[None].expand(n, m, *b.shape[1:]).reshape(n * m, *b.shape[1:]).contiguous()
    _, _, res, _ = kabsch_transformation_estimation(x1, x2)
    res_mat = res.mean(dim=1).reshape(n, m).contiguous()
    m0 = torch.empty(n, dtype=torch.int64, device=dev)
    m1 = torch.empty(m, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        ws = _pair_workspace(n, m, dev)
        rc = _lib.lib().ls_seq_match_scored(z0.data_ptr(), z1.data_ptr(), z0.shape[1], n, m, res_mat.data_ptr(), mode,
                                            m0.data_ptr(), m1.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "ls_seq_match_scored")
        _lib.launch_count += 1
        ws.record_stream(torch.cuda.current_stream(dev))
    return {"matches0": m0, "matches1": m1}


def sim3_seq_matcher(src_codes, tgt_codes):
    return _residual_seq_matcher(src_codes, tgt_codes, 1)


def eq_seq_matcher(src_codes, tgt_codes):
    return _residual_seq_matcher(src_codes, tgt_codes, 2)
