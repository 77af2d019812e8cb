"""Generate tests/golden/*.npz by executing the UNMODIFIED reference code.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference ships no tests or golden vectors, so parity is pinned
by running its own files on seeded inputs:

* ``arch_unet.py``, ``adapter.py``, ``utils_eval.py`` are imported as-is;
* ``train.py`` cannot be imported (module-level argparse / dataset glob / .cuda()),
  so ``space_to_depth`` / ``generate_mask_pair`` / ``generate_subimages`` are
  AST-extracted and executed unmodified, with ``get_generator`` taken from
  training_script.md:4-10 on a CPU generator;
* the N2N loop body is training_script.md:128-156;
* ``evaluation_704.py`` lines 57-68 and 74-120 are exec'd verbatim from the file.

The script also asserts that ``oracle/n2n_oracle.py`` reproduces every vector, so a
successful run == "oracle pinned against the reference".
"""
from __future__ import annotations

import ast
import os
import sys
import textwrap
from collections import OrderedDict

import numpy as np
import torch

REF = os.environ.get("N2N_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import n2n_oracle as O  # noqa: E402


def load_train_functions():
    src = open(os.path.join(REF, "train.py")).read()
    tree = ast.parse(src)
    want = {"space_to_depth", "generate_mask_pair", "generate_subimages"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {n.name for n in body} == want
    ns = {"torch": torch}
    state = {"operation_seed_counter": 0}

    def get_generator():                         # training_script.md:4-10, CPU device
        state["operation_seed_counter"] += 1
        g = torch.Generator(device="cpu")
        g.manual_seed(state["operation_seed_counter"])
        return g

    ns["get_generator"] = get_generator
    exec(compile(ast.Module(body=body, type_ignores=[]), "train.py", "exec"), ns)
    return ns, state


def csum(t: torch.Tensor):
    t = t.detach().double()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()], np.float64)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    import arch_unet
    import adapter as ref_adapter
    import utils_eval
    tf, counter = load_train_functions()

    # ------------------------------------------------------------------ sub-sampler
    sub = {}
    rng = np.random.RandomState(11)
    cases = [(2, 3, 8, 12), (1, 1, 32, 32), (3, 1, 6, 10)]
    for ci, (n, c, h, w) in enumerate(cases):
        img = torch.from_numpy(rng.randint(0, 1 << 20, size=(n, c, h, w)).astype(np.float32))
        seed_used = counter["operation_seed_counter"] + 1
        m1, m2 = tf["generate_mask_pair"](img)
        s1 = tf["generate_subimages"](img, m1)
        s2 = tf["generate_subimages"](img, m2)
        sub[f"img{ci}"] = img.numpy(); sub[f"seed{ci}"] = np.int64(seed_used)
        sub[f"m1_{ci}"] = m1.numpy(); sub[f"m2_{ci}"] = m2.numpy()
        sub[f"s1_{ci}"] = s1.numpy(); sub[f"s2_{ci}"] = s2.numpy()
        # oracle check
        rd = O.draw_rd_idx(n, h, w, seed_used)
        om1, om2 = O.masks_from_rd_idx(rd)
        assert np.array_equal(om1, m1.numpy()) and np.array_equal(om2, m2.numpy())
        assert np.array_equal(O.subimage_from_mask(img.numpy(), om1), s1.numpy())
        assert np.array_equal(O.subimage_from_mask(img.numpy(), om2), s2.numpy())
        sub[f"rd{ci}"] = rd
    # all eight constant selectors on an arange image (indices readable from the output)
    img = torch.arange(2 * 1 * 4 * 6, dtype=torch.float32).reshape(2, 1, 4, 6)
    for r in range(8):
        rd = np.full((2 * 2 * 3,), r, np.int64)
        om1, om2 = O.masks_from_rd_idx(rd)
        s1 = tf["generate_subimages"](img, torch.from_numpy(om1))
        s2 = tf["generate_subimages"](img, torch.from_numpy(om2))
        assert np.array_equal(O.subimage_from_mask(img.numpy(), om1), s1.numpy())
        sub[f"const_s1_{r}"] = s1.numpy(); sub[f"const_s2_{r}"] = s2.numpy()
    sub["const_img"] = img.numpy()
    np.savez_compressed(os.path.join(OUT, "subsample.npz"), **sub)

    # ------------------------------------------------------------------ UNet fwd + N2N step
    un = {}
    for tag, (in_nc, nf, hw, seed) in {"g1": (1, 4, 64, 3), "c3": (3, 4, 32, 5), "nf16": (1, 16, 32, 7)}.items():
        p = O.unet_init(in_nc, in_nc, nf, seed)
        net = arch_unet.UNet(in_nc=in_nc, out_nc=in_nc, n_feature=nf)
        assert list(net.state_dict().keys()) == list(p.keys())
        for k, v in net.state_dict().items():
            assert tuple(v.shape) == tuple(p[k].shape), k
        net.load_state_dict(p)
        # give biases non-zero values so bias paths are exercised
        g = torch.Generator().manual_seed(seed + 100)
        for k in p:
            if k.endswith(".bias"):
                p[k] = torch.randn(p[k].shape, generator=g) * 0.05
        net.load_state_dict(p)
        x = torch.rand(2, in_nc, hw, hw, generator=g)
        with torch.no_grad():
            y = net(x)
            yo = O.unet_forward(p, x)
        assert torch.allclose(y, yo, atol=1e-6, rtol=0), (y - yo).abs().max()
        un[f"{tag}_x"] = x.numpy(); un[f"{tag}_y"] = y.numpy()
        un[f"{tag}_bias_seed"] = np.int64(seed + 100)
        un[f"{tag}_wsum"] = np.stack([csum(v) for v in p.values()])
        if tag == "g1":
            # N2N step (training_script.md:128-156) with Lambda = epoch/n_epoch*ratio
            counter["operation_seed_counter"] = 40
            lam = 1 / 100 * 2.0
            clean = torch.rand(2, 1, hw, hw, generator=g)
            noisy = clean + torch.randn(clean.shape, generator=g) * (25.0 / 255.0)
            opt = torch.optim.Adam(net.parameters(), lr=3e-4)
            losses = []
            for it in range(3):
                opt.zero_grad()
                mask1, mask2 = tf["generate_mask_pair"](noisy)
                noisy_sub1 = tf["generate_subimages"](noisy, mask1)
                noisy_sub2 = tf["generate_subimages"](noisy, mask2)
                with torch.no_grad():
                    noisy_denoised = net(noisy)
                noisy_sub1_denoised = tf["generate_subimages"](noisy_denoised, mask1)
                noisy_sub2_denoised = tf["generate_subimages"](noisy_denoised, mask2)
                noisy_output = net(noisy_sub1)
                noisy_target = noisy_sub2
                Lambda = lam
                diff = noisy_output - noisy_target
                exp_diff = noisy_sub1_denoised - noisy_sub2_denoised
                loss1 = torch.mean(diff ** 2)
                loss2 = Lambda * torch.mean((diff - exp_diff) ** 2)
                loss_all = loss1 + loss2
                loss_all.backward()
                if it == 0:
                    un["step_noisy"] = noisy.numpy()
                    un["step_mask1"] = mask1.numpy(); un["step_mask2"] = mask2.numpy()
                    un["step_lambda"] = np.float64(lam)
                    un["step_loss"] = np.array([loss_all.item(), loss1.item(), loss2.item()])
                    for k, v in net.named_parameters():
                        un["grad/" + k] = v.grad.numpy().copy()
                    # oracle check
                    lo, l1o, l2o, go, _, _ = O.n2n_step_grads(p, noisy, mask1.numpy(), mask2.numpy(), lam)
                    assert abs(lo - loss_all.item()) < 1e-7
                    for k, v in net.named_parameters():
                        assert torch.allclose(go[k], v.grad, atol=1e-7, rtol=1e-5), k
                opt.step()
                losses.append(loss_all.item())
            un["step_losses3"] = np.array(losses)
            un["step_w3sum"] = np.stack([csum(v) for v in net.state_dict().values()])
            for k in ("enc_conv0.weight", "up3.deconv.weight", "nin_c.weight", "nin_c.bias", "dec_conv1a.bias"):
                un["w3/" + k] = net.state_dict()[k].numpy().copy()
    np.savez_compressed(os.path.join(OUT, "unet.npz"), **un)

    # ------------------------------------------------------------------ PSNR / SSIM
    ps = {}
    rng = np.random.RandomState(5)
    for i, shp in enumerate([(40, 56), (33, 47), (32, 32, 3), (24, 30, 1)]):
        a = rng.randint(0, 256, size=shp).astype(np.uint8)
        b = np.clip(a.astype(np.int32) + rng.randint(-20, 21, size=shp), 0, 255).astype(np.uint8)
        ps[f"a{i}"] = a; ps[f"b{i}"] = b
        ps[f"psnr{i}"] = np.float64(utils_eval.calculate_psnr(a, b))
        ps[f"ssim{i}"] = np.float64(utils_eval.calculate_ssim(a, b))
        assert abs(O.calculate_psnr(a, b) - ps[f"psnr{i}"]) < 1e-4
        assert abs(O.calculate_ssim(a, b) - ps[f"ssim{i}"]) < 1e-12, (O.calculate_ssim(a, b), ps[f"ssim{i}"])
    import cv2
    assert np.allclose(O.gaussian_taps(), cv2.getGaussianKernel(11, 1.5)[:, 0], atol=1e-15)
    np.savez_compressed(os.path.join(OUT, "psnr_ssim.npz"), **ps)

    # ------------------------------------------------------------------ adapter + finetune loss
    ad = {}
    C = 3
    base_p = O.unet_init(C, C, 4, 21)
    base = arch_unet.UNet(in_nc=C, out_nc=C, n_feature=4)
    base.load_state_dict(base_p)
    torch.manual_seed(9)
    model = ref_adapter.DenoiserWithAdapter(base, in_channels=C, hidden_channels=16)
    sd = model.state_dict()
    assert len(sd) == 50 + 4
    g = torch.Generator().manual_seed(77)
    clean = torch.rand(2, C, 32, 32, generator=g)
    noisy = clean + torch.randn(clean.shape, generator=g) * (25.0 / 255.0)
    pred = model(noisy)
    # finetune.py:283-285 via the reference's own definitions of gradient_loss
    ft_src = open(os.path.join(REF, "finetune.py")).read()
    ft_tree = ast.parse(ft_src)
    ft_body = [n for n in ft_tree.body if isinstance(n, ast.FunctionDef) and n.name in ("gradient", "gradient_loss")]
    ft_ns = {"torch": torch, "F": torch.nn.functional}
    exec(compile(ast.Module(body=ft_body, type_ignores=[]), "finetune.py", "exec"), ft_ns)
    loss_l1 = torch.nn.L1Loss()(pred, clean)
    loss_grad = ft_ns["gradient_loss"](pred, clean)
    loss = loss_l1 + 0.1 * loss_grad
    loss.backward()
    ad["noisy"] = noisy.numpy(); ad["clean"] = clean.numpy(); ad["pred"] = pred.detach().numpy()
    ad["loss"] = np.array([loss.item(), loss_l1.item(), loss_grad.item()])
    for k in ("adapter.net.0.weight", "adapter.net.0.bias", "adapter.net.2.weight", "adapter.net.2.bias"):
        ad["w/" + k] = sd[k].numpy().copy()
        ad["g/" + k] = dict(model.named_parameters())[k].grad.numpy().copy()
    ad["keys"] = np.array(list(sd.keys()))
    with torch.no_grad():
        bo = O.unet_forward(base_p, noisy)
    ap = {k: sd[k] for k in sd if k.startswith("adapter.")}
    po = O.adapter_forward(ap, noisy, bo)
    assert torch.allclose(po, pred, atol=1e-6)
    lo, _, _ = O.finetune_loss(po, clean, 0.1)
    assert abs(lo.item() - loss.item()) < 1e-7
    np.savez_compressed(os.path.join(OUT, "adapter.npz"), **ad)

    # ------------------------------------------------------------------ tiled 704-style eval
    ev = {}
    src_lines = open(os.path.join(REF, "evaluation_704.py")).read().split("\n")
    head = textwrap.dedent("\n".join(src_lines[56:68]))       # lines 57-68
    body = textwrap.dedent("\n".join(src_lines[73:120]))      # lines 74-120
    p = O.unet_init(1, 1, 4, 31)
    net = arch_unet.UNet(in_nc=1, out_nc=1, n_feature=4)
    net.load_state_dict(p); net.eval()
    from torchvision import transforms
    rng = np.random.RandomState(2025)
    H = W = 416
    field = rng.rand(H // 8 + 2, W // 8 + 2)
    clean_img = np.kron(field, np.ones((8, 8)))[:H, :W]
    clean_img = (255 * (clean_img - clean_img.min()) / (clean_img.max() - clean_img.min())).astype(np.float32)
    noisy_img = np.clip(clean_img + rng.randn(H, W) * 25.0, 0, 255).astype(np.float32)
    ns = {"np": np, "torch": torch, "network": net, "device": torch.device("cpu"),
          "transformer": transforms.Compose([transforms.ToTensor()]), "criterion": torch.nn.L1Loss(),
          "clean": clean_img, "noisy": noisy_img, "l1_list": []}
    exec(head, ns)
    exec(body, ns)
    pred255 = ns["pred255"]
    ev["noisy_u8"] = noisy_img.astype(np.uint8); ev["clean_u8"] = clean_img.astype(np.uint8)
    ev["pred255"] = pred255
    ev["psnr"] = np.float64(utils_eval.calculate_psnr(pred255, ns["clean"]))
    ev["ssim"] = np.float64(utils_eval.calculate_ssim(pred255, ns["clean"]))
    ev["weight_mask"] = ns["weight_mask"]
    assert np.array_equal(O.tile_weight(352), ns["weight_mask"])
    po = O.tiled_denoise(lambda t: O.unet_forward(p, t), noisy_img.astype(np.uint8))
    nd = int((po.astype(np.int32) != pred255.astype(np.int32)).sum())
    assert nd == 0, f"tiled oracle differs from reference at {nd} pixels"
    # whole-image (evaluation.py:73-83) on a crop
    x = torch.from_numpy(noisy_img[:64, :96] / 255.0)[None, None].float()
    with torch.no_grad():
        pr = net(x)
    prediction = pr.permute(0, 2, 3, 1).cpu().clamp(0, 1).numpy().squeeze()
    whole255 = np.clip(prediction * 255.0 + 0.5, 0, 255).astype(np.uint8)
    ev["whole255"] = whole255
    with torch.no_grad():
        assert np.array_equal(O.quantize_round(O.unet_forward(p, x)[0, 0].numpy()), whole255)
    np.savez_compressed(os.path.join(OUT, "eval.npz"), **ev)

    # ------------------------------------------------------------------ Adam / LR schedule
    am = {}
    g = torch.Generator().manual_seed(3)
    w0 = torch.randn(257, generator=g)
    w = torch.nn.Parameter(w0.clone())
    opt = torch.optim.Adam([w], lr=3e-4)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[int(20 * 0.1) - 1, int(40 * 0.1) - 1,
                                                                  int(60 * 0.1) - 1, int(80 * 0.1) - 1], gamma=0.5)
    grads = torch.randn(5, 257, generator=g)
    pw, pm, pv = w0.numpy().copy(), np.zeros(257, np.float32), np.zeros(257, np.float32)
    for t in range(5):
        w.grad = grads[t].clone()
        opt.step()
        O.adam_update(pw, grads[t].numpy(), pm, pv, t + 1, 3e-4)
    assert np.allclose(pw, w.detach().numpy(), atol=2e-7, rtol=0)
    am["w0"] = w0.numpy(); am["grads"] = grads.numpy(); am["w5"] = w.detach().numpy().copy()
    lrs = []
    for epoch in range(1, 11):
        lrs.append(opt.param_groups[0]["lr"])
        assert abs(O.multistep_lr(3e-4, epoch, 10, 0.5) - lrs[-1]) < 1e-12, (epoch, lrs[-1])
        sched.step()
    am["lrs_nepoch10"] = np.array(lrs)
    np.savez_compressed(os.path.join(OUT, "adam.npz"), **am)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    print("oracle pinned against reference: OK")


if __name__ == "__main__":
    main()
