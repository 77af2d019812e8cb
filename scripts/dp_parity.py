#!/usr/bin/env python
"""Launched under torchrun by tests/test_gpu_multirank.py (and usable by hand):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_parity.py

Every rank runs the data-parallel N2N step on its slice of a global batch; rank 0 also runs the same global batch
in one process; prints one JSON line with the gradient agreement and whether all replicas hold identical weights
after 5 optimiser steps (image_denoising_b200/selfcheck.py)."""
import datetime
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_denoising_b200 import N2NTrainer, UNet, selfcheck  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    rank, world = dist.get_rank(), dist.get_world_size()
    res = {}
    for precision in ("bf16", "fp32"):
        res[precision] = selfcheck.dp_gradient_parity(dev, per_rank=4, patch=128 if precision == "fp32" else 256,
                                                      precision=precision)
    # replicas stay identical over optimiser steps (graph replay path, per-rank data, global-batch masks)
    torch.manual_seed(5)
    net = UNet(1, 1, 48).to(dev)
    tr = N2NTrainer(net, lr=3e-4)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for it in range(5):
        clean = torch.rand((8, 1, 128, 128), generator=g, device=dev)
        tr.step(clean + torch.randn(clean.shape, generator=g, device=dev) * (25 / 255), 0.5)
    torch.cuda.synchronize()
    res["replicas_identical"] = selfcheck.replicas_identical(tr.flat_p)
    res["world"] = world
    if rank == 0:
        sys.stderr.flush()
        print("DP_PARITY " + json.dumps(res), flush=True)
    # the captured step graph holds NCCL work: drop it before the communicator goes away, and never let a stuck
    # teardown outlive the result that has already been printed
    del tr
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    import threading
    threading.Timer(30.0, lambda: os._exit(0)).start()
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    main()
