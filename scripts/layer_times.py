"""Per-launch CUDA-event times of the GEMM kernels for one bench-shaped N2N step (diagnostic)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import N2NTrainer, UNet, _ext

B = int(os.environ.get("B", "64"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(1, 1, 48).to(dev).set_precision("bf16")
tr = N2NTrainer(net, lr=3e-4, precision="bf16")
x = torch.rand(B, 1, 256, 256, device=dev)
for _ in range(3):
    tr.step(x, 0.02)
torch.cuda.synchronize()
L = _ext.lib()
L.n2n_profile_begin()
tr.step(x, 0.02)
buf = (ctypes.c_double * (3 * 400))()
n = L.n2n_profile_end_list(buf, 400)
names_fwd = ["enc0", "enc1", "enc2", "enc3", "enc4", "enc5", "enc6"] + ["up5"] * 4 + ["d5a", "d5b"] + ["up4"] * 4 + \
    ["d4a", "d4b"] + ["up3"] * 4 + ["d3a", "d3b"] + ["up2"] * 4 + ["d2a", "d2b"] + ["up1"] * 4 + ["d1a", "d1b", "nin_a", "nin_b", "nin_c"]
tot = {0: 0.0, 1: 0.0}
k = 0
for i in range(n):
    cls, ms, fl = int(buf[3 * i]), buf[3 * i + 1], buf[3 * i + 2]
    tot[cls] += ms
    tag = ""
    if cls == 0 and k < 2 * len(names_fwd):
        tag = ("full:" if k < len(names_fwd) else "half:") + names_fwd[k % len(names_fwd)]
        k += 1
    if ms > 0.0:
        print(f"{i:3d} cls={cls} {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TF/s(exec) {tag}")
print("totals ms:", tot)
