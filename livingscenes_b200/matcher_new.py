"""Instance matching on invariant codes -- drop-ins for lib_more/matcher_new.py.

  sequential_matcher(m0[n,256], m1[m,256])      matcher_new.py:109-139 (the matcher both evals select)
  nn_matcher(desc0[1,256,n], desc1[1,256,m])    matcher_new.py:85-105
Both return ``{'matches0', 'matches1'}`` int64 with -1 for unmatched, computed by one CTA per scene
pair on the GPU without any host synchronisation.  ``*_batched`` variants match many scene pairs in
one launch.  sinkhorn / sim3_seq / eq_seq matchers are not used by either eval.
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
