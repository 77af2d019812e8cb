import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from image_denoising_b200 import UNet
from oracle import n2n_oracle as O
dev = torch.device("cuda:0")
def weights(in_nc, nf, seed, scale=6.0):
    p = O.unet_init(in_nc, in_nc, nf, seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k in p:
        p[k] = torch.randn(p[k].shape, generator=g) * 0.05 if k.endswith(".bias") else p[k] * scale
    return p
def ref_intermediates(p, x):
    act = lambda t: F.leaky_relu(t, 0.2)
    c3 = lambda t, n: F.conv2d(t, p[n + ".weight"], p[n + ".bias"], padding=1)
    up = lambda t, n: F.conv_transpose2d(t, p[n + ".deconv.weight"], p[n + ".deconv.bias"], stride=2)
    r = {}
    skips = [x]
    t = act(c3(x, "enc_conv0")); r["enc_conv0"] = t
    t = act(c3(t, "enc_conv1")); r["enc_conv1"] = t; t = F.max_pool2d(t, 2); skips.append(t)
    for i in (2, 3, 4):
        t = act(c3(t, f"enc_conv{i}")); r[f"enc_conv{i}"] = t; t = F.max_pool2d(t, 2); skips.append(t)
    t = act(c3(t, "enc_conv5")); t = F.max_pool2d(t, 2); t = act(c3(t, "enc_conv6")); r["enc_conv6"] = t
    for lvl in (5, 4, 3, 2, 1):
        u = up(t, f"up{lvl}"); r[f"up{lvl}"] = u
        t = act(c3(torch.cat([u, skips[lvl - 1]], 1), f"dec_conv{lvl}a")); r[f"dec_conv{lvl}a"] = t
        t = act(c3(t, f"dec_conv{lvl}b")); r[f"dec_conv{lvl}b"] = t
    return r
for nf in (16, 32, 48):
    os.environ["N2N_NO_UPFUSE"] = "1"
    p = weights(1, nf, 11)
    net = UNet(1, 1, nf); net.load_state_dict(p); net = net.to(dev).set_precision("bf16")
    x = torch.rand(1, 1, 64, 64, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        y = net(x.to(dev)).cpu(); r = ref_intermediates(p, x)
    for name in ("enc_conv0", "enc_conv6", "dec_conv5a", "dec_conv5b", "dec_conv4a", "dec_conv3a", "dec_conv2a", "dec_conv2b", "dec_conv1a", "dec_conv1b"):
        try:
            got = net.read_activation(name).cpu()
        except Exception as e:
            print(name, "unavailable", e); continue
        ref = r[name]
        g = got[:, :ref.shape[1]]
        print(f"nf={nf} {name}: shape {tuple(got.shape)} nan={int(torch.isnan(g).sum())} err={(g-ref).abs().max().item():.3e} ref={ref.abs().max().item():.3e}")
    cat0 = net.read_activation("cat0").cpu()
    u = r["up1"]
    print(f"nf={nf} cat0 up-part: err={(cat0[:, :u.shape[1]]-u).abs().max().item():.3e} ref={u.abs().max().item():.3e}; skip block: {cat0[0, u.shape[1]:u.shape[1]+16, 5, 5].tolist()[:10]} x={x[0,0,4:7,4:7].flatten().tolist()}")
    d1a = net.read_activation("dec_conv1a").cpu(); ref = r["dec_conv1a"]
    bad = ((d1a[:, :96] - ref).abs() > 0.05 * ref.abs().max()) | torch.isnan(d1a[:, :96])
    print("bad d1a elements:", int(bad.sum()), "of", bad.numel(), "channels with bad:", bad.any(dim=(0, 2, 3)).nonzero().flatten().tolist()[:40],
          "rows:", bad.any(dim=(0, 1, 3)).nonzero().flatten().tolist()[:20], "cols:", bad.any(dim=(0, 1, 2)).nonzero().flatten().tolist()[:20])
