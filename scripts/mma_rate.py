"""tcgen05.mma issue-rate microbenchmark (diagnostic): cycles per M=128 x N x K=16 MMA by smem layout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import _ext
L = _ext.lib()
dev = torch.device("cuda:0")
iters = 4096
for nblocks in (1, 148):
    for layout, name in ((0, "SW32"), (2, "SW64"), (1, "SW128")):
        for n in (48, 96, 128, 192, 256):
            for nacc in (1, 2):
                if nacc * n > 512:
                    continue
                cyc = torch.zeros(nblocks, dtype=torch.int64, device=dev)
                _ext.check(L.n2n_probe_mma_rate(layout, n, iters, nacc, cyc.data_ptr(), nblocks, torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
                c = cyc.float().mean().item() / iters
                print(f"blocks={nblocks:3d} {name:5s} N={n:3d} accum={nacc}: {c:7.1f} cycles/MMA  (floor {128*n/256:.0f})")
