"""Data-parallel plumbing for the hot path: segments are independent from audio to tokens
(the reference's only partitioned axis is 'data', notebook :270-275), so each rank
(one process per GPU) takes a contiguous block of segments, the weights are broadcast ONCE at
load and the decoded token streams are all-gathered ONCE at the end.  No per-step collective.

Works with any torch.distributed backend: NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(num_segments: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ceil(S/N) segments for `rank` (keeps a file's segments in
    order for the stitch); trailing ranks may get fewer or none."""
    per = -(-num_segments // world_size) if num_segments > 0 else 0
    lo = min(num_segments, rank * per)
    return lo, min(num_segments, lo + per)


def broadcast_weights(blob: torch.Tensor, src: int = 0) -> torch.Tensor:
    """One broadcast of the flat fp32 weight blob (183.6 MB for mt3) from `src`."""
    _, n = world()
    if n > 1:
        dist.broadcast(blob, src=src)
    return blob


def broadcast_params(params, cfg, device, src: int = 0):
    """Weights at load, data-parallel: only rank `src` needs `params` ({tree path: array}); the flat
    blob is broadcast once and every rank gets the same {path: array} dict back.  With one process
    this is the identity."""
    from . import weights
    rank, n = world()
    if n == 1:
        return params
    if rank == src:
        blob = torch.from_numpy(weights.flatten(params, cfg)).to(device)
    else:
        blob = torch.empty(weights.num_params(cfg), dtype=torch.float32, device=device)
    broadcast_weights(blob, src=src)
    return weights.unflatten(blob.cpu().numpy(), cfg)


def gather_tokens(local_tokens: torch.Tensor, num_segments: int) -> torch.Tensor:
    """All-gather of the per-rank int32 [S_local, L] token streams -> [num_segments, L] on every
    rank, in global segment order.  Shards are padded to ceil(S/N) rows for the collective."""
    rank, n = world()
    if n == 1:
        return local_tokens[:num_segments]
    per = -(-num_segments // n)
    L = local_tokens.shape[1]
    padded = torch.zeros((per, L), dtype=local_tokens.dtype, device=local_tokens.device)
    padded[:local_tokens.shape[0]] = local_tokens
    out: List[torch.Tensor] = [torch.empty_like(padded) for _ in range(n)]
    dist.all_gather(out, padded)
    return torch.cat(out, dim=0)[:num_segments]
