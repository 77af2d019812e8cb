"""tcgen05.mma issue-rate microbenchmark (diagnostic): cycles per M=128 x N x K=16 MMA, same operands
every time vs eight distinct A / B tiles cycled (what a real kernel does)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import _ext
L = _ext.lib()
dev = torch.device("cuda:0")
iters = 4096
for nblocks in (1, 148):
    for layout, name in ((0, "SW32 same"), (8, "SW32 distinct")):
        for n in (48, 96, 128, 192, 256):
            for nacc in (1, 2):
                if nacc * n > 512 or (layout == 8 and n > 96):
                    continue
                cyc = torch.zeros(nblocks, dtype=torch.int64, device=dev)
                _ext.check(L.n2n_probe_mma_rate(layout, n, iters, nacc, cyc.data_ptr(), nblocks, torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
                c = cyc.float().mean().item() / iters
                print(f"blocks={nblocks:3d} {name:14s} N={n:3d} accum={nacc}: {c:7.1f} cycles/MMA  (math floor {128*n/256:.0f})")
