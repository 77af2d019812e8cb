import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import FusedAdam, ImprovedUNet, RESNET, Structure_loss, UNet
dev = torch.device("cuda:0")
def timed(fn, warm=3, it=6):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
out = {}
for name, ctor, batch, hw in (("unet", UNet, 16, 256), ("resnet", RESNET, 4, 256), ("improved", ImprovedUNet, 4, 128), ("improved256", ImprovedUNet, 4, 256)):
    torch.manual_seed(3)
    net = ctor(1, 1, 48).to(dev).set_precision("bf16")
    opt = FusedAdam(net.parameters(), lr=1e-4); crit = Structure_loss()
    clean = torch.rand(batch, 1, hw, hw, device=dev); noisy = clean + torch.randn_like(clean) * (25.0 / 255.0)
    def two():
        opt.zero_grad(); loss = crit(net(noisy), net(clean), clean); loss.backward(); opt.step()
    def fused():
        opt.zero_grad(); a, b = net(torch.cat([noisy, clean])).chunk(2); loss = crit(a, b, clean); loss.backward(); opt.step()
    out[name] = {"two_forwards_ms": timed(two), "one_batched_forward_ms": timed(fused)}
    del net, opt; torch.cuda.empty_cache()
print(json.dumps(out))
