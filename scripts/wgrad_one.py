"""One weight-gradient launch of a bench-shaped layer (diagnostic target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import ops
dev = torch.device("cuda:0")
n, cin, cout, h, w, k = [int(v) for v in os.environ.get("CASE", "64,16,96,128,128,1").split(",")]
x = torch.randn(n, cin, h, w, device=dev); dy = torch.randn(n, cout, h, w, device=dev)
for _ in range(3):
    ops.conv2d_wgrad(x, dy, k, "bf16")
torch.cuda.synchronize()
print("ok")
