"""Instance-sharded data parallelism over the GPUs of one box.

The reference has no distributed code; instances are independent through the encoder, and matching
needs every embedding of both scans, so the path needs exactly ONE exchange: an all-gather of the
packed per-instance records ``[z_so3 768 | z_inv 256 | s 1 | t 3]`` (1028 fp32 = 4112 B) over
NCCL / NVLink.  One process per GPU; rank r encodes the contiguous block
``[r*ceil(B/G), ...)`` of the instance list.  Works with the ``gloo`` backend on CPU tensors for the
host-logic tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

CODE_FLOATS = 1028


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition; the first ``n_items % world`` ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_codes(code: dict) -> torch.Tensor:
    """code dict -> [B,1028] records (the CUDA encoder can also emit these directly)."""
    B = code["z_inv"].shape[0]
    return torch.cat([code["z_so3"].reshape(B, -1), code["z_inv"], code["s"].reshape(B, 1),
                      code["t"].reshape(B, 3)], 1).contiguous()


def unpack_codes(rec: torch.Tensor, c_dim: int = 256) -> dict:
    B = rec.shape[0]
    return {"z_so3": rec[:, :3 * c_dim].reshape(B, c_dim, 3), "z_inv": rec[:, 3 * c_dim:4 * c_dim],
            "s": rec[:, 4 * c_dim], "t": rec[:, 4 * c_dim + 1:4 * c_dim + 4].reshape(B, 1, 3)}


def all_gather_codes(local_rec: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank [B_r,1028] records into the full [n_total,1028] table (same order as
    the un-sharded instance list).  Ragged shards are padded to the largest shard so that a single
    ``all_gather_into_tensor`` (NCCL) moves everything."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        assert local_rec.shape[0] == n_total
        return local_rec
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, local_rec.shape[1], dtype=local_rec.dtype, device=local_rec.device)
    pad[:local_rec.shape[0]] = local_rec
    out = torch.empty(world * mx, local_rec.shape[1], dtype=local_rec.dtype, device=local_rec.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(hi - lo == mx for lo, hi in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)


def encode_sharded(model, x_all: torch.Tensor, group=None) -> dict:
    """Encode this rank's block of ``x_all`` [B,3,N] (every rank holds, or can slice, the list) and
    all-gather the embeddings; returns the full code dict on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(x_all.shape[0], rank, world)
    rec = model.encode_packed(x_all[lo:hi])["packed"] if hi > lo else \
        torch.zeros(0, CODE_FLOATS, device=x_all.device)
    return unpack_codes(all_gather_codes(rec, x_all.shape[0], group))
