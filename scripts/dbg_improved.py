import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from image_denoising_b200 import ops, improved, ImprovedUNet, Structure_loss
from oracle import n2n_oracle as O
dev = torch.device("cuda:0")
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "r2_improved.npz"))
tag = "g16"
in_nc, nf, seed = (int(v) for v in z[f"{tag}_cfg"])
p = O.improved_init(in_nc, in_nc, nf, seed)
net = ImprovedUNet(in_nc=in_nc, out_nc=in_nc, n_feature=nf); net.load_state_dict(p); net = net.to(dev).set_precision("fp32")
noisy = torch.from_numpy(z[f"{tag}_noisy"]).to(dev); clean = torch.from_numpy(z[f"{tag}_clean"]).to(dev)
for mode in ("single-mse", "double-structure"):
    net.zero_grad()
    pr = {k: v.clone().to(dev).requires_grad_(True) for k, v in p.items()}
    prc = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    if mode == "single-mse":
        ((net(noisy) - clean) ** 2).mean().backward()
        ((O.improved_forward(pr, noisy) - clean) ** 2).mean().backward()
        ((O.improved_forward(prc, noisy.cpu()) - clean.cpu()) ** 2).mean().backward()
    else:
        Structure_loss()(net(noisy), net(clean), clean).backward()
        O.structure_loss(O.improved_forward(pr, noisy), O.improved_forward(pr, clean), clean)[0].backward()
        O.structure_loss(O.improved_forward(prc, noisy.cpu()), O.improved_forward(prc, clean.cpu()), clean.cpu())[0].backward()
    rows = []
    for k, v in net.named_parameters():
        a, b, c = v.grad, pr[k].grad, prc[k].grad.to(dev)
        s = b.abs().max().item()
        rows.append((k, (a - b).abs().max().item() / s, (a - c).abs().max().item() / s, (b - c).abs().max().item() / s))
    print(mode, "worst mine-vs-torchGPU %.1e  mine-vs-CPU %.1e  torchGPU-vs-CPU %.1e" % tuple(max(r[i] for r in rows) for i in (1, 2, 3)))
    for r in rows[-40:]:
        print("  %-30s %.1e %.1e %.1e" % r)
