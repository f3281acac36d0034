"""Stand-alone graph ops at the reference's pytorch3d boundary, executed by the CUDA library.

  knn_points              pytorch3d.ops.knn.knn_points as called at vec_dgcnn_atten.py:139-141
  sample_farthest_points  pytorch3d.ops.sample_farthest_points (vec_dgcnn_atten.py:169,
                          model_utils.py:205, more_solver.py:67,107-108)
  iterative_closest_point pytorch3d.ops.iterative_closest_point (more_solver.py:182-187)
"""
from __future__ import annotations

import collections
import ctypes as C

import torch

from . import _lib


@torch.no_grad()
def knn_points(p1: torch.Tensor, p2: torch.Tensor, K: int = 16, return_nn: bool = False):
    """p1 [B,P1,D], p2 [B,P2,D] -> (dists [B,P1,K] squared L2, idx [B,P1,K] int64, nn | None).
    Ascending distance; ties resolved towards the lower p2 index."""
    if K != _lib.LS_KNN_K:
        raise NotImplementedError("ls_knn is built for K == 16 (num_knn of the shipped model)")
    _lib.require_cuda(p1, "p1")
    B, P1, D = p1.shape
    P2 = p2.shape[1]
    q = p1.detach().float().transpose(1, 2).contiguous()  # [B,D,P1] channel-major
    s = p2.detach().float().transpose(1, 2).contiguous()
    idx = torch.empty(B, P1, K, dtype=torch.int64, device=p1.device)
    d2 = torch.empty(B, P1, K, dtype=torch.float32, device=p1.device)
    with torch.cuda.device(p1.device):
        rc = _lib.lib().ls_knn(q.data_ptr(), s.data_ptr(), B, D, P1, P2, idx.data_ptr(), d2.data_ptr(),
                               _lib.stream_ptr(p1.device))
        _lib.check(rc, "ls_knn")
        _lib.launch_count += 1
    nn = None
    if return_nn:
        nn = torch.gather(p2[:, None].expand(B, P1, P2, D), 2, idx[..., None].expand(B, P1, K, D))
    return d2, idx, nn


@torch.no_grad()
def knn_graph_cm(query_cm: torch.Tensor, source_cm: torch.Tensor):
    """Channel-major variant: query [B,D,Nq], source [B,D,Ns] -> (idx [B,Nq,16] int64, dist2)."""
    _lib.require_cuda(query_cm, "query")
    B, D, Nq = query_cm.shape
    Ns = source_cm.shape[2]
    q = query_cm.detach().float().contiguous()
    s = source_cm.detach().float().contiguous()
    idx = torch.empty(B, Nq, _lib.LS_KNN_K, dtype=torch.int64, device=q.device)
    d2 = torch.empty(B, Nq, _lib.LS_KNN_K, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        rc = _lib.lib().ls_knn(q.data_ptr(), s.data_ptr(), B, D, Nq, Ns, idx.data_ptr(), d2.data_ptr(),
                               _lib.stream_ptr(q.device))
        _lib.check(rc, "ls_knn")
        _lib.launch_count += 1
    return idx, d2


@torch.no_grad()
def knn_graph_cm_tc(query_cm: torch.Tensor, source_cm: torch.Tensor):
    """knn_graph_cm through the tensor-core candidate filter + exact re-rank (the encoder's path for Ns > 128).
    -> (idx [B,Nq,16] int64, dist2 [B,Nq,16], n_candidates [B,Nq] int32; -1 = overflow, brute-forced)."""
    _lib.require_cuda(query_cm, "query")
    B, D, Nq = query_cm.shape
    Ns = source_cm.shape[2]
    q = query_cm.detach().float().contiguous()
    s = source_cm.detach().float().contiguous()
    idx = torch.empty(B, Nq, _lib.LS_KNN_K, dtype=torch.int64, device=q.device)
    d2 = torch.empty(B, Nq, _lib.LS_KNN_K, dtype=torch.float32, device=q.device)
    nc = torch.empty(B, Nq, dtype=torch.int32, device=q.device)
    nbytes = C.c_size_t(0)
    _lib.check(_lib.lib().ls_knn_tc_workspace_bytes(B, D, Nq, Ns, C.byref(nbytes)), "ls_knn_tc_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        rc = _lib.lib().ls_knn_tc(q.data_ptr(), s.data_ptr(), B, D, Nq, Ns, idx.data_ptr(), d2.data_ptr(), nc.data_ptr(),
                                  ws.data_ptr(), ws.numel(), _lib.stream_ptr(q.device))
        _lib.check(rc, "ls_knn_tc")
        _lib.launch_count += 1
    return idx, d2, nc


@torch.no_grad()
def farthest_point_sample(xyz_cm: torch.Tensor, n_out: int, start_idx: torch.Tensor = None):
    """xyz [B,3,N] (channel-major) -> (idx [B,n_out] int64, sampled xyz [B,3,n_out]).  ``start_idx`` [B] int64
    chooses the first selected point per instance (default 0, pytorch3d's random_start_point=False)."""
    _lib.require_cuda(xyz_cm, "xyz")
    B, three, N = xyz_cm.shape
    assert three == 3
    x = xyz_cm.detach().float().contiguous()
    idx = torch.empty(B, n_out, dtype=torch.int64, device=x.device)
    out = torch.empty(B, 3, n_out, dtype=torch.float32, device=x.device)
    nbytes = C.c_size_t(0)
    _lib.check(_lib.lib().ls_fps_workspace_bytes(B, N, C.byref(nbytes)), "ls_fps_workspace_bytes")
    ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=x.device)
    st = None
    if start_idx is not None:
        st = start_idx.to(device=x.device, dtype=torch.int64).contiguous()
        assert st.shape == (B,) and int(st.min()) >= 0 and int(st.max()) < N
    with torch.cuda.device(x.device):
        rc = _lib.lib().ls_fps_ex(x.data_ptr(), B, N, n_out, _lib.ptr(st), idx.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                  ws.numel(), _lib.stream_ptr(x.device))
        _lib.check(rc, "ls_fps_ex")
        _lib.launch_count += 1
    return idx, out


@torch.no_grad()
def farthest_point_sample_masked(xyz_cm: torch.Tensor, mask: torch.Tensor, n_out: int, start_idx: torch.Tensor = None):
    """Batched FPS of ragged instances: xyz [B,3,Nmax], mask [B,Nmax] (or [B,1,Nmax]) bool -> (sampled xyz
    [B,3,n_out], n_valid [B] int32).  One launch for the whole batch; equivalent to FPS of ``pc[:, mask]`` per
    instance (model_utils.py:203-205)."""
    _lib.require_cuda(xyz_cm, "xyz")
    B, three, Nmax = xyz_cm.shape
    assert three == 3
    x = xyz_cm.detach().float().contiguous()
    m = mask.reshape(B, Nmax).to(device=x.device, dtype=torch.uint8).contiguous()
    out = torch.empty(B, 3, n_out, dtype=torch.float32, device=x.device)
    nv = torch.empty(B, dtype=torch.int32, device=x.device)
    ws = torch.empty(B * Nmax * 16, dtype=torch.uint8, device=x.device)
    st = None
    if start_idx is not None:
        st = start_idx.to(device=x.device, dtype=torch.int64).contiguous()
    with torch.cuda.device(x.device):
        rc = _lib.lib().ls_fps_masked(x.data_ptr(), m.data_ptr(), B, Nmax, n_out, _lib.ptr(st), None, out.data_ptr(),
                                      nv.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device))
        _lib.check(rc, "ls_fps_masked")
        _lib.launch_count += 1
    return out, nv


def sample_farthest_points(points: torch.Tensor, K: int = 50, random_start_point: bool = False):
    """pytorch3d signature: points [B,P,3] -> (pts [B,K,3], idx [B,K]).  ``random_start_point`` draws the first
    index per instance from torch's global RNG (pytorch3d draws it from its own RNG: the sequences differ, the
    distribution does not)."""
    start = None
    if random_start_point:
        start = torch.randint(0, points.shape[1], (points.shape[0],), dtype=torch.int64)
    idx, out = farthest_point_sample(points.transpose(1, 2), K, start)
    return out.transpose(1, 2).to(points.dtype), idx


SimilarityTransform = collections.namedtuple("SimilarityTransform", ["R", "T", "s"])
ICPSolution = collections.namedtuple("ICPSolution", ["converged", "rmse", "Xt", "RTs", "t_history"])


@torch.no_grad()
def iterative_closest_point(X: torch.Tensor, Y: torch.Tensor, init_transform=None, max_iterations: int = 100,
                            relative_rmse_thr: float = 1e-6, estimate_scale: bool = False,
                            allow_reflection: bool = False):
    """pytorch3d.ops.iterative_closest_point as called at more_solver.py:183 (tensors, equal lengths, rigid):
    X [B,N,3], Y [B,M,3], ``init_transform`` = (R [B,3,3], T [B,3], s [B]) in pytorch3d's row-vector convention
    (Xt = s X R + T).  Returns ICPSolution(converged, rmse [B], Xt [B,N,3], SimilarityTransform(R, T, s), []);
    ``t_history`` is not recorded.  One CTA per pair runs the whole loop (ls_icp)."""
    if estimate_scale or allow_reflection:
        raise NotImplementedError("ls_icp is built for the reference's call: estimate_scale=False, allow_reflection=False")
    _lib.require_cuda(X, "X")
    B, N, _ = X.shape
    M = Y.shape[1]
    x = X.detach().float().contiguous()
    y = Y.detach().float().contiguous()
    R0 = T0 = None
    if init_transform is not None:
        R0, T0, s0 = init_transform
        if not torch.all(s0 == 1):
            raise NotImplementedError("ls_icp: the initial scale must be 1 (rigid ICP)")
        R0 = R0.detach().float().expand(B, 3, 3).contiguous()
        T0 = T0.detach().float().expand(B, 3).contiguous()
    dev = x.device
    R = torch.empty(B, 3, 3, device=dev)
    T = torch.empty(B, 3, device=dev)
    rmse = torch.empty(B, device=dev)
    n_it = torch.empty(B, dtype=torch.int32, device=dev)
    Xt = torch.empty(B, N, 3, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().ls_icp(x.data_ptr(), y.data_ptr(), B, N, M, _lib.ptr(R0), _lib.ptr(T0), int(max_iterations),
                               float(relative_rmse_thr), R.data_ptr(), T.data_ptr(), rmse.data_ptr(), n_it.data_ptr(),
                               Xt.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, "ls_icp")
        _lib.launch_count += 1
    sol = ICPSolution(bool((n_it > 0).all()), rmse, Xt, SimilarityTransform(R, T, torch.ones(B, device=dev)), [])
    sol.RTs.R.n_iter = n_it  # iterations per pair (negative: not converged), for diagnostics
    return sol


@torch.no_grad()
def vn_linear(weight: torch.Tensor, v: torch.Tensor, tensor_cores: bool = True) -> torch.Tensor:
    """VecLinear.forward (vec_layers.py:121-134, so3, vector-only): weight [Co,Ci], v [B,Ci,3,...] ->
    [B,Co,3,...].  ``tensor_cores`` selects the tcgen05 3xTF32 GEMM (fp32-accurate) or the FP32 SIMT GEMM."""
    _lib.require_cuda(v, "v")
    B, Ci = v.shape[0], v.shape[1]
    Co = weight.shape[0]
    x = v.detach().float().reshape(B, Ci, -1).contiguous()
    n = x.shape[2]
    Kp = (Ci + 7) // 8 * 8
    w = torch.zeros(Co, Kp, device=v.device)
    w[:, :Ci] = weight.detach().float()
    out = torch.empty(B, Co, n, device=v.device)
    packed = _lib.tc_pack(w[:, :Ci].contiguous()) if tensor_cores else None
    with torch.cuda.device(v.device):
        rc = _lib.lib().ls_vn_linear(w.data_ptr(), _lib.ptr(packed), x.data_ptr(), out.data_ptr(), Co, Ci, Kp, B, n,
                                     _lib.stream_ptr(v.device))
        _lib.check(rc, "ls_vn_linear")
        _lib.launch_count += 1
    return out.reshape(B, Co, *v.shape[2:])
