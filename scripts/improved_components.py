import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import FusedAdam, ImprovedUNet, Structure_loss
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = ImprovedUNet(1, 1, 48).to(dev).set_precision("bf16")
opt = FusedAdam(net.parameters(), lr=1e-4); crit = Structure_loss()
clean = torch.rand(4, 1, 128, 128, device=dev); noisy = (clean + 0.1 * torch.randn_like(clean)).clamp(0, 1)
def timed(fn, k=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(k): fn()
    e1.record(); cpu = (time.perf_counter() - t0) / k * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, cpu
st = {}
def fwd(): st["a"], st["b"] = net(noisy), net(clean)
def fwd_loss_bwd():
    opt.zero_grad()
    loss = crit(net(noisy), net(clean), clean); loss.backward()
def full():
    opt.zero_grad()
    loss = crit(net(noisy), net(clean), clean); loss.backward(); opt.step()
print("two forwards (grad)   gpu %.2f ms cpu %.2f ms" % timed(fwd))
print("fwd + loss + bwd      gpu %.2f ms cpu %.2f ms" % timed(fwd_loss_bwd))
print("full step             gpu %.2f ms cpu %.2f ms" % timed(full))
with torch.no_grad():
    print("one no-grad forward   gpu %.2f ms cpu %.2f ms" % timed(lambda: net(noisy)))
