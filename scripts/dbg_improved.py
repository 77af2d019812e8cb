import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from image_denoising_b200 import ImprovedUNet, improved
from oracle import n2n_oracle as O
dev = torch.device("cuda:0")
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "r2_improved.npz"))
tag = "c48"
in_nc, nf, seed = (int(v) for v in z[f"{tag}_cfg"])
p = O.improved_init(in_nc, in_nc, nf, seed)
net = ImprovedUNet(in_nc=in_nc, out_nc=in_nc, n_feature=nf); net.load_state_dict(p); net = net.to(dev).set_precision("fp32")
noisy = torch.from_numpy(z[f"{tag}_noisy"]).to(dev); clean = torch.from_numpy(z[f"{tag}_clean"]).to(dev)
improved._DEBUG_TAPE = []
net.native_train = False
((net(noisy) - clean) ** 2).mean().backward()
tape = improved._DEBUG_TAPE; improved._DEBUG_TAPE = None
net.zero_grad(); net.native_train = True
((net(noisy) - clean) ** 2).mean().backward()
for lvl in range(4):
    for j, nm in enumerate(("t1", "t2", "t3")):
        t = tape[lvl][j]; idx = 9 + j + 6 * lvl
        a = net.read_buffer(idx, False); g = net.read_buffer(idx, True); c = t.shape[1]
        d = (g[:, :c] - t.grad)
        print(f"level {lvl} {nm}: act diff {(a[:, :c] - t.detach()).abs().max().item():.1e}  grad rel diff {d.abs().max().item() / t.grad.abs().max().item():.1e}"
              f"  worst channel {int(d.abs().amax(dim=(0, 2, 3)).argmax())} of {c}  per-channel err {[f'{v:.0e}' for v in (d.abs().amax(dim=(0,2,3)) / t.grad.abs().max()).tolist()[:: max(1, c // 12)]]}")
t = tape[2][0]; g = net.read_buffer(9 + 12, True)[:, :192]
d = ((g - t.grad).abs().amax(dim=(0, 2, 3)) / t.grad.abs().max()).tolist()
print("level 2 t1 per-channel rel err (channel: err) for err > 1e-4:", [(c, f"{v:.0e}") for c, v in enumerate(d) if v > 1e-4])
print("per-pixel err map of worst channel:", ((g - t.grad)[0, 152] / t.grad.abs().max()).cpu().numpy().round(3).tolist()[:3])
