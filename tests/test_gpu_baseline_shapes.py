"""Parity at the shapes BASELINE.json names (configs[0], [1], [3], [4]) — the CUDA path through the C-ABI
against the CPU oracle on the same seeded inputs, at the configs' own sizes:

* C1  N2N training, UNet(1,1,48), batch 4 x 1x256x256, sigma 25: loss, all 50 gradients, all weights after 1 and
      10 Adam steps (fp32 engine, max-abs); the bf16 engine against the same oracle run (10-step loss curve,
      per-tensor gradient cosine, weight drift);
* C2  generate_mask_pair + generate_subimages on 32 x 1x512x512 (fp32, bf16, and C = 3): torch.equal;
* C4  one 704x704 image through the tiled evaluation (evaluation_704.py semantics) at nf = 48;
* C5  adapter finetune, UNet(3,3,48) frozen, batch 4 x 3x256x256: loss, the 4 adapter gradients, adapter weights
      after 10 steps.
The oracle runs take ~1-2 s per step on the box's host cores."""
import numpy as np
import pytest
import torch

from oracle import n2n_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _net(dev, in_nc, nf, params, precision):
    from image_denoising_b200 import UNet
    net = UNet(in_nc=in_nc, out_nc=in_nc, n_feature=nf)
    net.load_state_dict(params)
    return net.to(dev).set_precision(precision)


def _weights(in_nc, nf, seed, bias_seed):
    p = O.unet_init(in_nc, in_nc, nf, seed)
    g = torch.Generator().manual_seed(int(bias_seed))
    for k in p:
        if k.endswith(".bias"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.05
    return p


# --------------------------------------------------------------------------- C1
@pytest.fixture(scope="module")
def c1_oracle():
    """BASELINE configs[0] on the oracle: 10 N2N + Adam steps, batch 4 x 1x256x256, sigma 25, Lambda = 0.02,
    lr 3e-4.  Selector seeds = operation counter 1..10."""
    torch.manual_seed(0)
    p0 = _weights(1, 48, 101, 102)
    clean = torch.rand(4, 1, 256, 256)
    noisy = O.add_train_noise_gauss(clean, 25.0, seed=1000)
    p = {k: v.clone() for k, v in p0.items()}
    m = {k: np.zeros(v.shape, np.float32) for k, v in p.items()}
    v_ = {k: np.zeros(v.shape, np.float32) for k, v in p.items()}
    lam, lr = 1 / 100 * 2.0, 3e-4
    losses, grads1, w1 = [], None, None
    rds = []
    for it in range(10):
        rd = O.draw_rd_idx(4, 256, 256, it + 1)
        rds.append(rd)
        m1, m2 = O.masks_from_rd_idx(rd)
        loss, l1, l2, grads, _, _ = O.n2n_step_grads(p, noisy, m1, m2, lam)
        losses.append([loss, l1, l2])
        if it == 0:
            grads1 = {k: g.numpy().copy() for k, g in grads.items()}
        for k in p:
            w = p[k].numpy()
            O.adam_update(w, grads[k].numpy(), m[k], v_[k], it + 1, lr)
        if it == 0:
            w1 = {k: t.numpy().copy() for k, t in p.items()}
    w10 = {k: t.numpy().copy() for k, t in p.items()}
    return dict(p0=p0, noisy=noisy, rds=rds, lam=lam, lr=lr, losses=np.array(losses), grads1=grads1, w1=w1, w10=w10)


def _run_trainer(dev, c1, precision):
    from image_denoising_b200 import N2NTrainer
    net = _net(dev, 1, 48, c1["p0"], precision)
    tr = N2NTrainer(net, lr=c1["lr"], precision=precision)
    noisy = c1["noisy"].to(dev)
    losses, grads1, w1 = [], None, None
    for it in range(10):
        l3 = tr.step(noisy, c1["lam"], rd_idx=torch.from_numpy(c1["rds"][it]).to(dev))
        losses.append(l3.cpu().numpy().copy())
        if it == 0:
            grads1 = [g.cpu().numpy().copy() for g in tr.grads]
            w1 = {k: v.cpu().numpy().copy() for k, v in net.state_dict().items()}
    w10 = {k: v.cpu().numpy().copy() for k, v in net.state_dict().items()}
    return np.array(losses), grads1, w1, w10


def test_c1_train_step_fp32_loss_grads_weights_1_and_10_steps(dev, c1_oracle):
    c1 = c1_oracle
    losses, grads1, w1, w10 = _run_trainer(dev, c1, "fp32")
    assert np.allclose(losses, c1["losses"], rtol=2e-5, atol=1e-9), (losses[:, 0], c1["losses"][:, 0])
    worst = 0.0
    for (k, ref), got in zip(c1["grads1"].items(), grads1):
        ref = ref.astype(np.float64); got = got.astype(np.float64)
        rel = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
        worst = max(worst, rel)
        assert rel <= 3e-4, (k, rel)                       # fp32 summation-order noise over 4x128x128 pixels
    for tag, ref_w, got_w in (("1 step", c1["w1"], w1), ("10 steps", c1["w10"], w10)):
        for k in ref_w:
            # Adam's first steps move every weight by ~lr regardless of gradient scale: compare against that
            # step size (a sign flip on a ~zero gradient entry costs at most 2*lr per step)
            err = np.abs(got_w[k] - ref_w[k]).max()
            assert err <= (2e-5 if tag == "1 step" else 3e-4), (tag, k, err)
        drift = np.mean([np.abs(got_w[k] - ref_w[k]).mean() for k in ref_w])
        assert drift <= 2e-6, (tag, drift)
    print(f"C1 fp32: worst relative gradient error {worst:.3e}; loss curve max rel "
          f"{np.abs(losses[:, 0] / c1['losses'][:, 0] - 1).max():.2e}")


def test_c1_train_bf16_ten_step_loss_curve_and_grads(dev, c1_oracle):
    c1 = c1_oracle
    losses, grads1, w1, w10 = _run_trainer(dev, c1, "bf16")
    rel = np.abs(losses[:, 0] / c1["losses"][:, 0] - 1)
    assert rel.max() <= 2e-2, rel                          # every point of the 10-step loss curve within 2 %
    assert losses[-1, 0] < losses[0, 0] and c1["losses"][-1, 0] < c1["losses"][0, 0]
    bad = []
    for (k, ref), got in zip(c1["grads1"].items(), grads1):
        a, b = got.astype(np.float64).ravel(), ref.astype(np.float64).ravel()
        cos = float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))
        if cos < 0.99:
            bad.append((k, cos))
    assert not bad, bad
    # weights after 10 Adam steps: the mean displacement from the fp32 oracle stays a small fraction of the
    # 10*lr = 3e-3 each weight can have moved
    drift = np.mean([np.abs(w10[k] - c1["w10"][k]).mean() for k in w10])
    assert drift <= 3e-4, drift
    print(f"C1 bf16: loss curve rel err max {rel.max():.3e}, mean weight drift after 10 steps {drift:.2e}")


# --------------------------------------------------------------------------- C2
@pytest.mark.parametrize("dtype,c", [(torch.float32, 1), (torch.bfloat16, 1), (torch.float32, 3)])
def test_c2_subsampler_32x512x512_bit_exact(dev, dtype, c):
    """BASELINE configs[1]: generate_mask_pair + generate_subimages on 32 x c x 512 x 512, every index checked."""
    from image_denoising_b200 import generate_subimage_pair, generate_subimages, n2n, ops
    n, h, w = 32, 512, 512
    g = torch.Generator().manual_seed(c * 10 + (dtype == torch.bfloat16))
    img = torch.rand(n, c, h, w, generator=g).to(dtype)
    for seed in (1, 16):
        rd = O.draw_rd_idx(n, h, w, seed)
        m1, m2 = O.masks_from_rd_idx(rd)
        img_np = img.view(torch.int16).numpy() if dtype == torch.bfloat16 else img.numpy()
        ref1, ref2 = O.subimage_from_mask(img_np, m1), O.subimage_from_mask(img_np, m2)
        d_m1, d_m2, pk = ops.mask_pair_from_rdidx(torch.from_numpy(rd).to(dev), want_masks=True, want_packed=True)
        assert torch.equal(d_m1.cpu(), torch.from_numpy(m1)) and torch.equal(d_m2.cpu(), torch.from_numpy(m2))
        x = img.to(dev)
        s1 = generate_subimages(x, d_m1); s2 = generate_subimages(x, d_m2)
        p1, p2 = generate_subimage_pair(x, packed=pk)
        q1, q2 = generate_subimage_pair(x, d_m1, d_m2)
        for got, ref in ((s1, ref1), (s2, ref2), (p1, ref1), (p2, ref2), (q1, ref1), (q2, ref2)):
            assert got.dtype == dtype and got.shape == (n, c, h // 2, w // 2)
            got_np = got.cpu().view(torch.int16).numpy() if dtype == torch.bfloat16 else got.cpu().numpy()
            assert np.array_equal(got_np, ref)
    # the reference call surface: generate_mask_pair(img) draws from get_generator() on the image's device
    n2n.operation_seed_counter = 0
    a1, a2 = n2n.generate_mask_pair(img.to(dev))
    assert n2n.operation_seed_counter == 1 and a1.dtype == torch.bool and a1.numel() == n * h // 2 * w // 2 * 4
    k = a1.view(-1, 4).int().sum(1)
    assert bool((k == 1).all()) and bool((a2.view(-1, 4).int().sum(1) == 1).all()) and not bool((a1 & a2).any())


def test_space_to_depth_matches_oracle(dev):
    """train.py:134-138 (public name; never materialised by the fused sub-sampler)."""
    from image_denoising_b200 import space_to_depth
    g = torch.Generator().manual_seed(3)
    for shape, bs in (((2, 3, 8, 12), 2), ((1, 1, 512, 512), 2), ((2, 2, 12, 18), 3), ((1, 5, 4, 4), 1)):
        x = torch.rand(shape, generator=g)
        ref = torch.nn.functional.unfold(x, bs, stride=bs).view(shape[0], shape[1] * bs * bs, shape[2] // bs, shape[3] // bs)
        assert np.array_equal(O.space_to_depth(x.numpy(), bs), ref.numpy())
        assert torch.equal(space_to_depth(x.to(dev), bs).cpu(), ref)
    with pytest.raises(ValueError):
        space_to_depth(torch.zeros(1, 1, 5, 4, device=dev), 2)


# --------------------------------------------------------------------------- C4
def test_c4_tiled_704_image_vs_oracle_nf48(dev):
    """One 704x704 grayscale image through evaluation_704.py's 9-tile path at nf = 48: fp32 engine within 1 LSB on
    a vanishing fraction of pixels; bf16 engine within PSNR 0.01 dB of the oracle's output (north-star tolerance)."""
    from image_denoising_b200 import evaluate, utils_eval
    p = _weights(1, 48, 31, 32)
    rng = np.random.RandomState(2025)
    field = rng.rand(704 // 8 + 2, 704 // 8 + 2)
    clean = np.kron(field, np.ones((8, 8)))[:704, :704]
    clean = (255 * (clean - clean.min()) / (clean.max() - clean.min())).astype(np.uint8)
    noisy = np.clip(clean.astype(np.float32) + rng.randn(704, 704) * 25.0, 0, 255).astype(np.uint8)
    ref = O.tiled_denoise(lambda t: O.unet_forward(p, t), noisy)
    assert (ref[0, :] == 0).all() and (ref[:, 0] == 0).all()
    out32, _ = evaluate.denoise_tiled(_net(dev, 1, 48, p, "fp32"), [noisy])
    d = np.abs(out32[0].astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3, (d.max(), (d > 0).mean())
    out16, _ = evaluate.denoise_tiled(_net(dev, 1, 48, p, "bf16"), [noisy])
    assert (out16[0][0, :] == 0).all() and (out16[0][:, 0] == 0).all()
    assert abs(utils_eval.calculate_psnr(out16[0], clean) - O.calculate_psnr(ref, clean)) < 0.01
    assert abs(utils_eval.calculate_ssim(out16[0], clean) - O.calculate_ssim(ref, clean)) < 1e-4
    assert np.abs(out16[0].astype(np.int32) - ref.astype(np.int32)).max() <= 3


# --------------------------------------------------------------------------- C5
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_c5_adapter_finetune_4x3x256x256_nf48(dev, precision):
    """BASELINE configs[4] at the CPU-runnable batch 4: frozen UNet(3,3,48) + OutputAdapter(3, 16), loss
    L1 + 0.1*gradient loss, Adam lr 1e-4 over the adapter only; 10 steps."""
    from image_denoising_b200 import DenoiserWithAdapter, FusedAdam, UNet, l1_grad_loss
    base_p = _weights(3, 48, 7, 8)
    g = torch.Generator().manual_seed(7)
    clean = torch.rand(4, 3, 256, 256, generator=g)
    noisy = clean + torch.randn(clean.shape, generator=g) * (25 / 255)
    ap = {"adapter.net.0.weight": (torch.rand(16, 6, 3, 3, generator=g) - 0.5) * 0.27,
          "adapter.net.0.bias": (torch.rand(16, generator=g) - 0.5) * 0.27,
          "adapter.net.2.weight": (torch.rand(3, 16, 3, 3, generator=g) - 0.5) * 0.16,
          "adapter.net.2.bias": (torch.rand(3, generator=g) - 0.5) * 0.16}
    # oracle: the base is frozen, so its output is computed once
    with torch.no_grad():
        base_out = O.unet_forward(base_p, noisy)
    ref_p = {k: v.clone().requires_grad_(True) for k, v in ap.items()}
    opt_ref = torch.optim.Adam(list(ref_p.values()), lr=1e-4)
    ref_losses, ref_g1 = [], None
    for it in range(10):
        opt_ref.zero_grad()
        loss, l1, lg = O.finetune_loss(O.adapter_forward(ref_p, noisy, base_out), clean, 0.1)
        loss.backward()
        if it == 0:
            ref_g1 = {k: v.grad.clone() for k, v in ref_p.items()}
        opt_ref.step()
        ref_losses.append([loss.item(), l1.item(), lg.item()])
    base = UNet(in_nc=3, out_nc=3, n_feature=48)
    base.load_state_dict(base_p)
    model = DenoiserWithAdapter(base, in_channels=3, hidden_channels=16)
    sd = model.state_dict(); sd.update(ap); model.load_state_dict(sd)
    model = model.to(dev).set_precision(precision)
    opt = FusedAdam(filter(lambda q: q.requires_grad, model.parameters()), lr=1e-4)
    x, c = noisy.to(dev), clean.to(dev)
    losses, g1 = [], None
    for it in range(10):
        opt.zero_grad(set_to_none=True)
        loss, loss3 = l1_grad_loss(model(x), c, 0.1)
        loss.backward()
        if it == 0:
            g1 = {k: q.grad.cpu().clone() for k, q in model.named_parameters() if q.grad is not None}
        opt.step()
        losses.append(loss3.cpu().numpy().copy())
    losses, ref_losses = np.array(losses), np.array(ref_losses)
    assert set(g1.keys()) == set(ap.keys())
    if precision == "fp32":
        assert np.allclose(losses, ref_losses, rtol=2e-5), (losses[:, 0], ref_losses[:, 0])
        for k in ap:
            ref = ref_g1[k]
            assert float((g1[k] - ref).abs().max()) <= 3e-4 * float(ref.abs().max()) + 1e-9, k
            assert float((model.state_dict()[k].cpu() - ref_p[k].detach()).abs().max()) <= 1e-4, k    # <= 10*lr*small
    else:
        assert np.abs(losses[:, 0] / ref_losses[:, 0] - 1).max() <= 2e-2
        for k in ap:
            a, b = g1[k].double().flatten(), ref_g1[k].double().flatten()
            assert float((a * b).sum() / (a.norm() * b.norm())) > 0.99, k
            assert float((model.state_dict()[k].cpu() - ref_p[k].detach()).abs().mean()) <= 3e-4, k
    assert losses[-1, 0] < losses[0, 0]
