"""ImprovedUNet timing on one B200: forward images/s and the supervised training step of train.py:354-368,
this repo's kernels vs stock PyTorch (cuDNN, fp32 and bf16 autocast) running the pinned oracle graph on the same GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import FusedAdam, ImprovedUNet, Structure_loss
from oracle import n2n_oracle as O

dev = torch.device("cuda:0")


def timeit(fn, warm=2, it=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(it):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it


out = {}
torch.manual_seed(0)
net = ImprovedUNet(1, 1, 48).to(dev)
p = {k: v.detach().clone() for k, v in net.state_dict().items()}
for prec in ("bf16", "fp32"):
    net.set_precision(prec)
    for n, hw in ((8, 256), (1, 704)):
        x = torch.rand(n, 1, hw, hw, device=dev)
        with torch.no_grad():
            ms = timeit(lambda: net(x), it=3 if prec == "fp32" else 5)
        out[f"fwd_{prec}_{n}x{hw}_ms"] = ms
x = torch.rand(8, 1, 256, 256, device=dev)
with torch.no_grad():
    out["torch_fp32_fwd_8x256_ms"] = timeit(lambda: O.improved_forward(p, x))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out["torch_bf16_fwd_8x256_ms"] = timeit(lambda: O.improved_forward(p, x))
net.set_precision("bf16")
opt = FusedAdam(net.parameters(), lr=1e-4); crit = Structure_loss()
clean = torch.rand(4, 1, 128, 128, device=dev); noisy = (clean + 0.1 * torch.randn_like(clean)).clamp(0, 1)


def step():
    opt.zero_grad()
    loss = crit(net(noisy), net(clean), clean)
    loss.backward()
    opt.step()


out["train_bf16_4x128_ms"] = timeit(step)
out["launches_fwd"], out["launches_bwd"] = net.last_launches, getattr(net, "last_bwd_launches", None)
clean_b, noisy_b = clean, noisy
clean = torch.rand(4, 1, 256, 256, device=dev); noisy = (clean + 0.1 * torch.randn_like(clean)).clamp(0, 1)
out["train_bf16_4x256_ms"] = timeit(step)
clean, noisy = clean_b, noisy_b
pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
topt = torch.optim.Adam(pr.values(), lr=1e-4)


def tstep():
    topt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a, b = O.improved_forward(pr, noisy), O.improved_forward(pr, clean)
    O.structure_loss(a.float(), b.float(), clean)[0].backward()
    topt.step()


out["torch_bf16_train_4x128_ms"] = timeit(tstep)
print(json.dumps(out))
