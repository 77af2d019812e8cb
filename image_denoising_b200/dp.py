"""Host-side data-parallel plumbing: batch sharding, gradient buckets and the
bucketed all-reduce.  Pure torch.distributed — works on the NCCL backend (GPU ranks) and on gloo
(the CPU tests of the host logic); none of it touches the kernels.

The reference's only multi-GPU mode is single-process ``nn.DataParallel`` (train.py:324-325):
scatter the batch, replicate the weights every step, reduce gradients onto GPU 0.  Here each GPU is
its own process; rank r owns samples [r*B/W, (r+1)*B/W) of the global batch, every op up to the loss
is per-sample and the loss is a mean, so with equal shards

    grad_global = (1/W) * sum_r grad_r

— one all-reduce(sum) of the flat fp32 gradient per step, the 1/W folded into the Adam kernel."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(rank: int, world: int, n: int) -> Tuple[int, int]:
    """Samples [lo, hi) of a global batch of ``n`` owned by ``rank``; shards must be equal so that
    the mean-of-means equals the global mean."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return rank * per, (rank + 1) * per


def shard_selector(rd_idx_global: torch.Tensor, rank: int, world: int, n: int) -> torch.Tensor:
    """This rank's slice of a selector drawn for the GLOBAL batch (cells are ordered (n, i, j),
    train.py:144-162), so that a W-rank run uses the masks a 1-GPU run would."""
    cells_per_sample = rd_idx_global.numel() // n
    lo, hi = shard_range(rank, world, n)
    return rd_idx_global[lo * cells_per_sample:hi * cells_per_sample]


def bucket_slices(sizes: Sequence[int], buckets: int = 2, tail_tensors: int = 10) -> List[Tuple[int, int]]:
    """Element ranges of the flat gradient buffer, in the order their gradients become ready:
    the head / full-resolution decoder tensors sit at the END of the state_dict order
    (arch_unet.py:177-190) and are produced first by the backward pass."""
    total = int(sum(sizes))
    if buckets <= 1 or len(sizes) <= tail_tensors:
        return [(0, total)]
    cut = total - int(sum(sizes[-tail_tensors:]))
    return [(cut, total), (0, cut)] if 0 < cut < total else [(0, total)]


def allreduce_buckets(flat_grad: torch.Tensor, slices: Sequence[Tuple[int, int]], group=None) -> None:
    """SUM all-reduce of each bucket (the 1/world average is applied by the optimiser kernel)."""
    for a, b in slices:
        dist.all_reduce(flat_grad[a:b], op=dist.ReduceOp.SUM, group=group)


def broadcast_params(flat_params: torch.Tensor, src: int = 0, group=None) -> None:
    """Once at start-up, instead of DataParallel's per-step replicate."""
    dist.broadcast(flat_params, src=src, group=group)


def allreduce_mean_grads(params, group=None) -> None:
    """Average the ``.grad`` of ``params`` over the ranks with ONE all-reduce of a flat buffer (the autograd training loops of
    entry/train.py: a per-parameter all-reduce is 50 .. 240 latency-bound collectives per step).  Parameters without a
    gradient (RESNET's unused up5, arch_unet.py:303) are skipped on every rank alike."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    world = dist.get_world_size(group)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
