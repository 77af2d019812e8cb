"""Adapter finetune step (BASELINE configs[4]: frozen base UNet(3,3,48) + OutputAdapter, batch 32x3x256x256) —
throughput of the public-API step (model forward, fused L1+gradient loss, backward, FusedAdam)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import DenoiserWithAdapter, FusedAdam, UNet, l1_grad_loss
dev = torch.device("cuda:0")
torch.manual_seed(7)
base = UNet(3, 3, 48)
model = DenoiserWithAdapter(base, in_channels=3, hidden_channels=16).to(dev)
model.base.set_precision("bf16")
opt = FusedAdam([p for p in model.parameters() if p.requires_grad], lr=1e-4)
clean = torch.rand(32, 3, 256, 256, device=dev)
noisy = clean + torch.randn_like(clean) * (25 / 255)
def step():
    opt.zero_grad(set_to_none=True)
    loss, _ = l1_grad_loss(model(noisy), clean, 0.1)
    loss.backward()
    opt.step()
    return loss
for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 20
for _ in range(K):
    l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"adapter finetune step 32x3x256x256 bf16: {ms:.2f} ms/step -> {32 / ms * 1e3:.0f} patches/s, {32 * 39.45 / ms:.0f} TFLOP/s algorithmic, loss {float(l):.5f}")
