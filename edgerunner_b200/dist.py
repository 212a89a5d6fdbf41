"""Multi-GPU host logic.  One process per GPU (torchrun), `torch.distributed` for the plumbing.

* Decode does not shard (one request is one sequential chain; ``LMM.generate`` asserts B == 1): requests are dealt out
  to ranks round-robin and each rank runs an independent replica — no collective on the data path.  ``gather_tokens``
  brings the generated id arrays back to rank 0 for reporting only.
* The teacher-forced forward shards by batch (data parallel).  Its one collective is an all-reduce of the
  (sum of per-token losses, token count, KL sum) triple, which reproduces the single-process mean exactly
  (``F.cross_entropy(..., ignore_index=-100)`` is a mean over the non-ignored tokens of the WHOLE batch).
* ``FlatGradAllReduce``: the collective a data-parallel TRAINING step needs (one flattened gradient buffer,
  sum then / world, in a few slices in reverse-layer order so that it can start while earlier layers are still in
  backward).  ``edgerunner_b200.train.FlatTrainer`` exports the gradients of ``er_train_step`` straight into its buffer (``bench.py --workload train``
  under torchrun); ``bench.py --workload tf --grad-allreduce`` times it on a synthetic buffer next to the forward-only workload, the gloo test covers the plumbing.

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


class FlatGradAllReduce:
    """One flat fp32 buffer for all gradients (what DDP / accelerate build for the reference, main.py:133-142), all-reduced in ``n_slices``
    contiguous slices from the END of the buffer to its start (parameters are registered input-to-output, backward produces them
    output-to-input), asynchronously; ``wait()`` completes them and divides by the world size."""

    def __init__(self, numel: int, device, n_slices: int = 4, dtype=torch.float32):
        self.buf = torch.zeros(numel, dtype=dtype, device=device)
        self.n_slices = max(1, int(n_slices))
        self.handles = []
        step = -(-numel // self.n_slices)
        self.bounds = [(max(0, numel - (i + 1) * step), numel - i * step) for i in range(self.n_slices) if numel - i * step > 0]

    def views(self, shapes):
        """Per-parameter views into the flat buffer, in registration order (what a backward pass would write its gradients into)."""
        out, off = [], 0
        for shp in shapes:
            n = int(np.prod(shp)) if len(shp) else 1
            out.append(self.buf[off:off + n].view(shp))
            off += n
        assert off <= self.buf.numel()
        return out

    def launch(self):
        rank, ws = world()
        self.handles = []
        if ws > 1:
            for lo, hi in self.bounds:
                self.handles.append(dist.all_reduce(self.buf[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        return self

    def wait(self):
        rank, ws = world()
        for h in self.handles:
            h.wait()
        self.handles = []
        if ws > 1:
            self.buf.div_(ws)
        return self.buf
