"""Launch-floor probe: the captured N2N step at small patch sizes (every layer is launch / prologue bound), so
ms_per_step / launches estimates the in-graph cost of one small launch — what the 8x8..32x32 levels of the 256x256 step pay.
argv: list of batch:hw pairs (default 64:64 16:64 64:128 64:256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import N2NTrainer, UNet
dev = torch.device("cuda:0")
shapes = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(64, 64), (16, 64), (64, 128), (64, 256)]
for b, hw in shapes:
    torch.manual_seed(0)
    net = UNet(1, 1, 48).to(dev)
    tr = N2NTrainer(net, lr=3e-4, precision="bf16")
    x = torch.rand(b, 1, hw, hw, device=dev)
    for _ in range(6):
        tr.step(x, 1.0)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 30
        for _ in range(K):
            tr.step(x, 1.0)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K)
    print(f"batch {b} x {hw}x{hw}: {best:.3f} ms/step, {tr.last_launches} launches -> {best * 1e3 / tr.last_launches:.2f} us per launch")
    del tr, net
