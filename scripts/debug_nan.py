import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import UNet, ops
from oracle import n2n_oracle as O
dev = torch.device("cuda:0")
def weights(in_nc, nf, seed, scale=6.0):
    p = O.unet_init(in_nc, in_nc, nf, seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k in p:
        p[k] = torch.randn(p[k].shape, generator=g) * 0.05 if k.endswith(".bias") else p[k] * scale
    return p
for nf in (16, 32, 48):
    for shape in ((3, 32, 64), (1, 64, 64), (2, 128, 128), (1, 32, 32)):
        for fuse in ("0", "1"):
            os.environ["N2N_NO_UPFUSE"] = fuse
            p = weights(1, nf, 11)
            net = UNet(1, 1, nf); net.load_state_dict(p); net = net.to(dev).set_precision("bf16")
            x = torch.rand(shape[0], 1, shape[1], shape[2], generator=torch.Generator().manual_seed(5))
            with torch.no_grad():
                y = net(x.to(dev)).cpu(); ref = O.unet_forward(p, x)
            print(f"nf={nf} shape={shape} NO_UPFUSE={fuse}: nan={int(torch.isnan(y).sum())} err={(y-ref).abs().max().item():.3e} ref={ref.abs().max().item():.3e} launches={net.last_launches}")
# single-layer deconv checks
for (ci, co, h, w) in ((32, 32, 4, 8), (32, 32, 2, 4), (16, 16, 1, 2), (32, 32, 8, 16), (96, 96, 4, 8)):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, ci, h, w, generator=g); wt = torch.randn(ci, co, 2, 2, generator=g) * 0.1; b = torch.randn(co, generator=g)
    y = ops.deconv2x2_fwd(x.to(dev), wt.to(dev), b.to(dev), precision="bf16").cpu()
    ref = torch.nn.functional.conv_transpose2d(x, wt, b, stride=2)
    print(f"deconv ci={ci} co={co} {h}x{w}: nan={int(torch.isnan(y).sum())} err={(y-ref).abs().max().item():.3e}")
for (ci, co, h, w) in ((32, 32, 4, 8), (48, 32, 4, 8), (32, 32, 2, 4), (16, 16, 2, 4), (16, 16, 1, 2), (32, 32, 8, 16)):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, ci, h, w, generator=g); wt = torch.randn(co, ci, 3, 3, generator=g) * 0.1; b = torch.randn(co, generator=g)
    y = ops.conv2d_fwd(x.to(dev), wt.to(dev), b.to(dev), 0.2, precision="bf16").cpu()
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, wt, b, padding=1), 0.2)
    print(f"conv ci={ci} co={co} {h}x{w}: nan={int(torch.isnan(y).sum())} err={(y-ref).abs().max().item():.3e}")
