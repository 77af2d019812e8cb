"""Per-launch CUDA-event times of the GEMM-class kernels for one bench-shaped N2N step (diagnostic).
Labels follow the launch structure of unet_plan.cu for the bf16 engine at 256x256 / 128x128."""
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


def fwd_names(res, fused=False):
    """launches of one forward at input resolution `res` (levels res .. res/32).  fused: the no-grad pass, where
    ConvTranspose + dec_conv a run as two fused launches (one per output-row parity) and no deconv launch exists."""
    def up(name, in_res):
        if fused and in_res >= 4 and os.environ.get("N2N_NO_UPFUSE", "0") != "1":
            return []
        return [name] * (2 if in_res >= 4 else 4)               # pair form (slab engine) from 4x4 inputs up
    def dxa(name, r):
        if fused and r // 2 >= 4 and os.environ.get("N2N_NO_UPFUSE", "0") != "1":
            return [name + "+up"] * 2
        return [name]                                           # CTA-pair engine: half the weights per SM, no K split
    r = res
    out = ["enc1", "enc2", "enc3", "enc4", "enc5", "enc6"]          # enc0 runs in the fused input stage (not a GEMM launch)
    out += up("up5", r // 32) + dxa("d5a", r // 16) + ["d5b"]
    out += up("up4", r // 16) + dxa("d4a", r // 8) + ["d4b"]
    out += up("up3", r // 8) + dxa("d3a", r // 4) + ["d3b"]
    out += up("up2", r // 4) + dxa("d2a", r // 2) + ["d2b"]
    out += up("up1", r // 2) + dxa("d1a", r) + ["d1b", "head"]
    return out


names = ["full:" + s for s in fwd_names(256, True)] + ["half:" + s for s in fwd_names(128)]
tot = {0: 0.0, 1: 0.0}
k = 0
for i in range(n):
    cls, ms, fl = int(buf[3 * i]), buf[3 * i + 1], buf[3 * i + 2]
    tot[cls] += ms
    tag = "wgrad" if cls == 1 else ("bwd" if k >= len(names) else names[k])
    if cls == 0:
        k += 1
    print(f"{i:3d} cls={cls} {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TF/s(exec) {tag}")
print("totals ms:", tot)
