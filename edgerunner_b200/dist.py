"""Multi-GPU host logic.  One process per GPU (torchrun), `torch.distributed` for the plumbing.

* Decode does not shard (one request is one sequential chain; ``LMM.generate`` asserts B == 1): requests are dealt out
  to ranks round-robin and each rank runs an independent replica — no collective on the data path.  ``gather_tokens``
  brings the generated id arrays back to rank 0 for reporting only.
* The teacher-forced forward shards by batch (data parallel).  Its one collective is an all-reduce of the
  (sum of per-token losses, token count, KL sum) triple, which reproduces the single-process mean exactly
  (``F.cross_entropy(..., ignore_index=-100)`` is a mean over the non-ignored tokens of the WHOLE batch).

Works with backend 'nccl' on the GPU box and 'gloo' on CPU (tests/test_dist_cpu.py, world_size 2).
"""

from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def shard_requests(n_requests: int) -> List[int]:
    """Indices of the requests this rank serves (round-robin: rank r takes r, r+W, r+2W, ...)."""
    rank, ws = world()
    return list(range(rank, n_requests, ws))


def max_over_ranks(x: float, device=None) -> float:
    """Device-timed durations are reported as the max over ranks (the job is as slow as its slowest replica)."""
    rank, ws = world()
    if ws == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_tokens(local: Sequence[np.ndarray]) -> List[np.ndarray] | None:
    """All ranks' generated id arrays on rank 0, in request order (inverse of ``shard_requests``); None elsewhere."""
    rank, ws = world()
    if ws == 1:
        return list(local)
    bucket = [None] * ws if rank == 0 else None
    dist.gather_object([np.asarray(a) for a in local], bucket, dst=0)
    if rank != 0:
        return None
    n = sum(len(b) for b in bucket)
    out: List[np.ndarray] = [None] * n
    for r, arrs in enumerate(bucket):
        for j, a in enumerate(arrs):
            out[r + j * ws] = a
    return out


def dp_reduce_losses(ce_sum: torch.Tensor, n_tokens: torch.Tensor, kl_sum: torch.Tensor, kl_weight: float):
    """Data-parallel loss of ``LMM.forward`` (core/models.py:189-199): ONE all-reduce of [ce_sum, n_tokens, kl_sum].

    Per rank: ce_sum = sum of token cross-entropies of its shard, n_tokens = number of supervised tokens, kl_sum =
    0.5 * sum(latent^2) of its shard.  Returns (loss, loss_ce, loss_kl) identical on every rank and equal to the
    single-process result on the concatenated batch."""
    buf = torch.stack([ce_sum.reshape(()).double(), n_tokens.reshape(()).double(), kl_sum.reshape(()).double()])
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    loss_ce = buf[0] / buf[1].clamp(min=1)
    loss_kl = buf[2]
    w = float(np.float32(kl_weight))          # the C ABI takes kl_weight as fp32: same constant as the single-process kernel
    return (loss_ce + w * loss_kl).float(), loss_ce.float(), loss_kl.float()
