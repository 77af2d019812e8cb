"""No-grad full-resolution UNet forward (bench shape) — a short target for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import UNet
B = int(os.environ.get("B", "32"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(1, 1, 48).to(dev).set_precision("bf16")
x = torch.rand(B, 1, 256, 256, device=dev)
with torch.no_grad():
    for _ in range(int(os.environ.get("REPS", "3"))):
        y = net(x)
torch.cuda.synchronize()
print("ok", float(y.mean()))
