import os, sys
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
for _ in range(3):
    opt.zero_grad(set_to_none=True)
    loss, _ = l1_grad_loss(model(noisy), clean, 0.1)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print(float(loss))
