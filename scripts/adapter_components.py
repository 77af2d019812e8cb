"""Where the C5 adapter-finetune step spends its time: CUDA-event time of each component, back to back (GPU time) and
the wall time of the whole public-API step (CPU launch overhead shows up as wall > sum)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import DenoiserWithAdapter, FusedAdam, UNet, l1_grad_loss
dev = torch.device("cuda:0")
torch.manual_seed(7)
model = DenoiserWithAdapter(UNet(3, 3, 48), in_channels=3, hidden_channels=16).to(dev)
model.set_precision("bf16")
opt = FusedAdam([p for p in model.parameters() if p.requires_grad], lr=1e-4)
clean = torch.rand(32, 3, 256, 256, device=dev)
noisy = clean + torch.randn_like(clean) * (25 / 255)


def timed(fn, k=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(k):
        fn()
    e1.record(); t_launch = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, t_launch / k * 1e3


with torch.no_grad():
    print("base forward          gpu %.3f ms  cpu-launch %.3f ms" % timed(lambda: model.base(noisy)))
    base_out = model.base(noisy)
    print("adapter forward nograd gpu %.3f ms  cpu-launch %.3f ms" % timed(lambda: model.adapter(noisy, base_out)))
state = {}
def fwd_only():
    state["pred"] = model.adapter(noisy, base_out)
print("adapter forward (grad) gpu %.3f ms  cpu-launch %.3f ms" % timed(fwd_only))
def loss_only():
    state["loss"], _ = l1_grad_loss(state["pred"].detach().requires_grad_(True), clean, 0.1)
print("loss fwd               gpu %.3f ms  cpu-launch %.3f ms" % timed(loss_only))
def fwd_bwd():
    opt.zero_grad(set_to_none=True)
    pred = model.adapter(noisy, base_out)
    loss, _ = l1_grad_loss(pred, clean, 0.1)
    loss.backward()
print("adapter fwd+loss+bwd   gpu %.3f ms  cpu-launch %.3f ms" % timed(fwd_bwd))
def full():
    opt.zero_grad(set_to_none=True)
    loss, _ = l1_grad_loss(model(noisy), clean, 0.1)
    loss.backward()
    opt.step()
print("full step              gpu %.3f ms  cpu-launch %.3f ms" % timed(full))
