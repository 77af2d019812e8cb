"""World-size-2 gloo test (CPU) of the data-parallel host logic: two ranks each run
the N2N step on their shard of a global batch (the oracle stands in for the kernels — this tests the
sharding / bucketing / all-reduce arithmetic, not the CUDA path), all-reduce the flat gradient in
buckets and apply the 1/world scale; the result must equal the 1-process global-batch gradient, and
both ranks must end with identical weights after an Adam step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import n2n_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    p = O.unet_init(1, 1, 4, 5)
    g = torch.Generator().manual_seed(11)
    clean = torch.rand(4, 1, 64, 64, generator=g)
    noisy = clean + torch.randn(clean.shape, generator=g) * (25 / 255)
    rd = O.draw_rd_idx(4, 64, 64, 1)
    return p, noisy, rd


def _worker(rank, world, port, out):
    from image_denoising_b200 import dp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        p, noisy, rd = _problem()
        n = noisy.shape[0]
        lo, hi = dp.shard_range(rank, world, n)
        rd_r = dp.shard_selector(torch.from_numpy(rd), rank, world, n).numpy()
        m1, m2 = O.masks_from_rd_idx(rd_r)
        loss, _, _, grads, _, _ = O.n2n_step_grads(p, noisy[lo:hi], m1, m2, 0.5)
        names = list(p.keys())
        sizes = [p[k].numel() for k in names]
        flat_p = torch.cat([p[k].reshape(-1) for k in names]).clone()
        if rank != 0:
            flat_p.zero_()                                    # only rank 0 holds the weights before the broadcast
        dp.broadcast_params(flat_p, 0)
        flat_g = torch.cat([grads[k].reshape(-1) for k in names]).clone()
        slices = dp.bucket_slices(sizes, buckets=2)
        assert sum(b - a for a, b in slices) == flat_g.numel() and len(slices) == 2
        dp.allreduce_buckets(flat_g, slices)
        flat_g *= 1.0 / world                                 # what the Adam kernel's grad_scale does
        m = np.zeros(flat_p.numel(), np.float32); v = np.zeros_like(m)
        w = flat_p.numpy().copy()
        O.adam_update(w, flat_g.numpy(), m, v, 1, 3e-4)
        # the autograd loops' helper: one flat all-reduce, mean over ranks, parameters without a gradient skipped
        prm = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(4))]
        prm[0].grad = torch.full((3, 2), float(rank + 1)); prm[2].grad = torch.arange(4.0) * (rank + 1)
        dp.allreduce_mean_grads(prm)
        assert torch.equal(prm[0].grad, torch.full((3, 2), 1.5)) and prm[1].grad is None
        assert torch.equal(prm[2].grad, torch.arange(4.0) * 1.5)
        out[rank] = (flat_g.numpy().copy(), w, float(loss))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_equals_global_batch():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    p, noisy, rd = _problem()
    m1, m2 = O.masks_from_rd_idx(rd)
    loss, _, _, grads, _, _ = O.n2n_step_grads(p, noisy, m1, m2, 0.5)
    ref = torch.cat([grads[k].reshape(-1) for k in p.keys()]).numpy()
    g0, w0, l0 = out[0]
    g1, w1, l1 = out[1]
    assert np.array_equal(g0, g1), "ranks disagree after the all-reduce"
    assert np.array_equal(w0, w1), "ranks diverged after the optimiser step"
    assert np.abs(g0 - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-30) + 1e-9
    assert abs(0.5 * (l0 + l1) - float(loss)) <= 1e-6 * abs(float(loss))


def test_shard_helpers():
    from image_denoising_b200 import dp
    assert dp.shard_range(1, 4, 64) == (16, 32)
    with pytest.raises(ValueError):
        dp.shard_range(0, 3, 64)
    rd = torch.arange(8 * 4)                                    # 8 samples x 4 cells
    assert dp.shard_selector(rd, 3, 4, 8).tolist() == list(range(24, 32))
    assert dp.bucket_slices([5, 5, 5], buckets=1) == [(0, 15)]
    sl = dp.bucket_slices([1] * 50, buckets=2)
    assert sl == [(40, 50), (0, 40)]
