"""One forward conv launch of a bench-shaped layer (diagnostic target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import ops
dev = torch.device("cuda:0")
n, cin, cout, h, w, k = [int(v) for v in os.environ.get("CASE", "64,48,48,256,256,3").split(",")]
x = torch.randn(n, cin, h, w, device=dev); wt = torch.randn(cout, cin, k, k, device=dev) * 0.05
b = torch.zeros(cout, device=dev)
for _ in range(3):
    ops.conv2d_fwd(x, wt, b, 0.2, "bf16")
torch.cuda.synchronize()
print("ok")
