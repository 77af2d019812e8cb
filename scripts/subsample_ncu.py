"""C2-shaped sub-sampler launches (32 x 1 x 512 x 512 fp32) — a short target for `ncu --set full -k regex:subsample_vec`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
imgs = [torch.rand((32, 1, 512, 512), generator=g, device=dev) for _ in range(6)]
rds = [torch.randint(0, 8, (32 * 256 * 256,), generator=g, device=dev) for _ in range(6)]
for i in range(12):
    m1, m2, pk = ops.mask_pair_from_rdidx(rds[i % 6], want_masks=True, want_packed=True)
    a, b = ops.subsample_pair(imgs[i % 6], m1, m2)
    c, d = ops.subsample_pair(imgs[i % 6], packed=pk)
torch.cuda.synchronize()
print("ok", float(a.sum() + c.sum()))
