"""On-hardware data-parallel parity checks (SURVEY.md §4 tier 4, §8e), shared by ``bench.py`` (the
``dp_parity`` key of its JSON line) and ``tests/test_gpu_multirank.py`` (launched under torchrun).

* ``dp_gradient_parity`` — W ranks x B patches reproduce the gradient one process computes on the
  concatenated W*B batch with the same weights, noise and masks (the selector is drawn for the global batch
  and sliced, ``dp.shard_selector``): per-sample arithmetic is identical, only the fp32 summation order of
  the weight gradients differs.
* ``replicas_identical`` — every rank holds bit-identical weights (exact integer checksum of the fp32 bit
  patterns) after the steps it has run.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import dp, n2n
from .arch_unet import UNet
from .trainer import N2NTrainer


def replicas_identical(flat_p: torch.Tensor, group=None) -> bool:
    """Exact: the int64 sum of the fp32 bit patterns plus the float64 sum must agree on every rank."""
    bits = flat_p.view(torch.int32).to(torch.int64).sum()
    pos = (flat_p.view(torch.int32).to(torch.int64) * torch.arange(1, flat_p.numel() + 1, device=flat_p.device)).sum()
    mine = torch.stack([bits, pos])
    world = dist.get_world_size(group)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    return all(torch.equal(v, allv[0]) for v in allv)


def dp_gradient_parity(dev, per_rank: int = 8, patch: int = 256, nf: int = 48, precision: str = "bf16", group=None):
    """Returns a dict {max_rel, cos, loss_rel, ok} (identical on every rank).  ok: max-abs gradient difference
    <= 1e-3 of the largest gradient entry, cosine >= 0.99999, mean-of-rank losses == global loss to 1e-5."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = per_rank * world
    torch.manual_seed(4242)                                             # same weights on every rank
    net = UNet(in_nc=1, out_nc=1, n_feature=nf).to(dev).set_precision(precision)
    state = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator(device=dev).manual_seed(99)                     # same global batch on every rank
    clean = torch.rand((n, 1, patch, patch), generator=g, device=dev)
    noisy = clean + torch.randn(clean.shape, generator=g, device=dev) * (25.0 / 255.0)
    saved_counter = n2n.operation_seed_counter
    n2n.operation_seed_counter = 777
    rd_global = n2n.draw_rd_idx(noisy)                                  # train.py:155-162 on the global batch
    n2n.operation_seed_counter = 777
    lo, hi = dp.shard_range(rank, world, n)
    tr = N2NTrainer(net, lr=0.0, precision=precision, process_group=group, use_graph=False)
    loss_r = tr.step(noisy[lo:hi], 0.5).clone()                         # draws the global selector itself and slices it
    n2n.operation_seed_counter = saved_counter
    g_dp = tr.flat_g.clone() / world                                    # all-reduced SUM -> mean (Adam applies 1/world)
    loss_sum = loss_r[0].clone()
    dist.all_reduce(loss_sum, group=group)
    res = torch.zeros(4, dtype=torch.float64, device=dev)
    if rank == 0:
        ref = UNet(in_nc=1, out_nc=1, n_feature=nf).to(dev).set_precision(precision)
        ref.load_state_dict(state)
        tr1 = N2NTrainer(ref, lr=0.0, precision=precision, use_graph=False, data_parallel=False)
        loss_g = tr1.step(noisy, 0.5, rd_idx=rd_global).clone()
        a, b = g_dp.double(), tr1.flat_g.double()
        res[0] = (a - b).abs().max() / b.abs().max()
        res[1] = (a * b).sum() / (a.norm() * b.norm())
        res[2] = ((loss_sum / world - loss_g[0]).abs() / loss_g[0].abs()).double()
        res[3] = 1.0 if (res[0] <= 1e-3 and res[1] >= 0.99999 and res[2] <= 1e-5) else 0.0
        del tr1, ref
    dist.broadcast(res, 0, group=group)
    del tr
    torch.cuda.empty_cache()
    return {"max_rel": float(res[0]), "cos": float(res[1]), "loss_rel": float(res[2]), "ok": bool(res[3] > 0),
            "per_rank": per_rank, "global_batch": n, "precision": precision}
