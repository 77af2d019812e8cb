"""CPU: the oracle (oracle/n2n_oracle.py) against the golden vectors produced by the
unmodified reference code (oracle/make_golden.py)."""
import numpy as np
import torch

from oracle import n2n_oracle as O


def _weights(in_nc, nf, seed, bias_seed):
    p = O.unet_init(in_nc, in_nc, nf, seed)
    g = torch.Generator().manual_seed(int(bias_seed))
    for k in p:
        if k.endswith(".bias"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.05
    return p, g


def _csum(t):
    t = t.detach().double()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def test_subsampler_matches_reference(golden):
    z = golden("subsample")
    for ci in range(3):
        img = z[f"img{ci}"]
        n, c, h, w = img.shape
        rd = O.draw_rd_idx(n, h, w, int(z[f"seed{ci}"]))
        assert np.array_equal(rd, z[f"rd{ci}"])
        m1, m2 = O.masks_from_rd_idx(rd)
        assert np.array_equal(m1, z[f"m1_{ci}"]) and np.array_equal(m2, z[f"m2_{ci}"])
        assert np.array_equal(O.subimage_from_mask(img, m1), z[f"s1_{ci}"])
        assert np.array_equal(O.subimage_from_mask(img, m2), z[f"s2_{ci}"])
    img = z["const_img"]
    for r in range(8):
        m1, m2 = O.masks_from_rd_idx(np.full((12,), r))
        assert np.array_equal(O.subimage_from_mask(img, m1), z[f"const_s1_{r}"])
        assert np.array_equal(O.subimage_from_mask(img, m2), z[f"const_s2_{r}"])


def test_subsampler_invariants():
    rng = np.random.RandomState(0)
    rd = rng.randint(0, 8, size=2 * 5 * 7)
    m1, m2 = O.masks_from_rd_idx(rd)
    a = m1.reshape(-1, 4); b = m2.reshape(-1, 4)
    assert (a.sum(1) == 1).all() and (b.sum(1) == 1).all()
    k1, k2 = a.argmax(1), b.argmax(1)
    assert (k1 != k2).all()
    # 4-adjacent, never diagonal: exactly one of (ky, kx) differs
    assert (((k1 // 2) != (k2 // 2)) ^ ((k1 % 2) != (k2 % 2))).all()


def test_unet_forward_matches_reference(golden):
    z = golden("unet")
    for tag, (in_nc, nf, seed) in {"g1": (1, 4, 3), "c3": (3, 4, 5), "nf16": (1, 16, 7)}.items():
        p, _ = _weights(in_nc, nf, seed, z[f"{tag}_bias_seed"])
        ws = np.stack([_csum(v) for v in p.values()])
        assert np.allclose(ws, z[f"{tag}_wsum"], rtol=1e-12), "seeded weights differ from fixture"
        with torch.no_grad():
            y = O.unet_forward(p, torch.from_numpy(z[f"{tag}_x"]))
        assert np.abs(y.numpy() - z[f"{tag}_y"]).max() < 1e-6


def test_n2n_step_matches_reference(golden):
    z = golden("unet")
    p, _ = _weights(1, 4, 3, z["g1_bias_seed"])
    loss, l1, l2, grads, _, _ = O.n2n_step_grads(
        p, torch.from_numpy(z["step_noisy"]), z["step_mask1"], z["step_mask2"], float(z["step_lambda"]))
    assert np.allclose([loss, l1, l2], z["step_loss"], atol=1e-7)
    for k, g in grads.items():
        ref = z["grad/" + k]
        assert np.abs(g.numpy() - ref).max() <= 1e-7 + 1e-5 * np.abs(ref).max(), k


def test_adam_and_lr_schedule(golden):
    z = golden("adam")
    w = z["w0"].copy(); m = np.zeros_like(w); v = np.zeros_like(w)
    for t in range(5):
        O.adam_update(w, z["grads"][t], m, v, t + 1, 3e-4)
    assert np.abs(w - z["w5"]).max() < 2e-7
    lrs = [O.multistep_lr(3e-4, e, 10, 0.5) for e in range(1, 11)]
    assert np.allclose(lrs, z["lrs_nepoch10"], rtol=1e-12)


def test_psnr_ssim_match_reference(golden):
    z = golden("psnr_ssim")
    for i in range(4):
        assert abs(O.calculate_psnr(z[f"a{i}"], z[f"b{i}"]) - float(z[f"psnr{i}"])) < 1e-4
        assert abs(O.calculate_ssim(z[f"a{i}"], z[f"b{i}"]) - float(z[f"ssim{i}"])) < 1e-12


def test_adapter_and_finetune_loss_match_reference(golden):
    z = golden("adapter")
    base_p = O.unet_init(3, 3, 4, 21)
    noisy = torch.from_numpy(z["noisy"]); clean = torch.from_numpy(z["clean"])
    ap = {k[2:]: torch.from_numpy(z[k]).requires_grad_(True) for k in z.files if k.startswith("w/")}
    with torch.no_grad():
        bo = O.unet_forward(base_p, noisy)
    pred = O.adapter_forward(ap, noisy, bo)
    assert np.abs(pred.detach().numpy() - z["pred"]).max() < 1e-6
    loss, l1, lg = O.finetune_loss(pred, clean, 0.1)
    assert np.allclose([loss.item(), l1.item(), lg.item()], z["loss"], atol=1e-7)
    loss.backward()
    for k, t in ap.items():
        ref = z["g/" + k]
        assert np.abs(t.grad.numpy() - ref).max() <= 1e-7 + 1e-5 * np.abs(ref).max(), k
    assert len(z["keys"]) == 54


def test_tiled_and_whole_eval_match_reference(golden):
    z = golden("eval")
    p = O.unet_init(1, 1, 4, 31)
    assert np.array_equal(O.tile_weight(352), z["weight_mask"])
    out = O.tiled_denoise(lambda t: O.unet_forward(p, t), z["noisy_u8"])
    assert np.array_equal(out, z["pred255"])
    assert (out[0, :] == 0).all() and (out[:, 0] == 0).all()
    assert abs(O.calculate_psnr(out, z["clean_u8"].astype(np.float32)) - float(z["psnr"])) < 1e-4
    assert abs(O.calculate_ssim(out, z["clean_u8"].astype(np.float32)) - float(z["ssim"])) < 1e-10
    x = torch.from_numpy(z["noisy_u8"][:64, :96].astype(np.float32) / 255.0)[None, None]
    with torch.no_grad():
        w = O.quantize_round(O.unet_forward(p, x)[0, 0].numpy())
    assert np.array_equal(w, z["whole255"])


def test_oracle_ref_modules_agree_with_the_restatement():
    """oracle/_ref (the reference's own modules staged by oracle/build_ref.py; present in the build container and on
    the GPU box, absent from a bare checkout) against the restatement: UNet forward and the sub-sampler."""
    import os
    import sys
    import pytest
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "train_functions.py")):
        pytest.skip("oracle/_ref not staged")
    sys.path.insert(0, ref)
    try:
        import arch_unet as ref_arch
        import train_functions as tf
    finally:
        sys.path.remove(ref)
    p = O.unet_init(1, 1, 8, 3)
    net = ref_arch.UNet(in_nc=1, out_nc=1, n_feature=8)
    net.load_state_dict(p)
    x = torch.rand(1, 1, 32, 64, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert torch.allclose(net(x), O.unet_forward(p, x), atol=1e-6)
    tf.operation_seed_counter = 6
    m1, m2 = tf.generate_mask_pair(x)
    o1, o2 = O.masks_from_rd_idx(O.draw_rd_idx(1, 32, 64, 7))
    assert np.array_equal(m1.numpy(), o1) and np.array_equal(m2.numpy(), o2)
    assert np.array_equal(tf.generate_subimages(x, m1).numpy(), O.subimage_from_mask(x.numpy(), o1))
    assert np.array_equal(tf.space_to_depth(x, 2).numpy(), O.space_to_depth(x.numpy(), 2))


def _rand_bias(p, seed):
    g = torch.Generator().manual_seed(seed)
    for k in p:
        if k.endswith(".bias"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.05
    return p


def test_space_to_depth_and_structure_loss_match_reference(golden):
    z = golden("r2_misc")
    for i in range(4):
        assert np.array_equal(O.space_to_depth(z[f"s2d_x{i}"], int(z[f"s2d_bs{i}"])), z[f"s2d_y{i}"])
    for i in range(3):
        pred = torch.from_numpy(z[f"sl_pred{i}"]).requires_grad_(True)
        pred2 = torch.from_numpy(z[f"sl_pred2{i}"]).requires_grad_(True)
        loss, px, tv, cs = O.structure_loss(pred, pred2, torch.from_numpy(z[f"sl_tgt{i}"]))
        loss.backward()
        assert np.allclose([loss.item(), px.item(), tv.item(), cs.item()], z[f"sl_loss{i}"], rtol=1e-6)
        assert np.allclose(pred.grad.numpy(), z[f"sl_g1_{i}"], atol=1e-8) and np.allclose(pred2.grad.numpy(), z[f"sl_g2_{i}"], atol=1e-8)


def test_iqsl_loss_matches_reference(golden):
    """finetune_iqsl.py:291-383, golden made by executing the reference's own function (oracle/make_golden_r2.py)."""
    z = golden("r2_misc")
    for i in range(3):
        t1, t2, tau, margin, cef = (float(v) for v in z[f"iq_cfg{i}"])
        pred = torch.from_numpy(z[f"iq_pred{i}"]).requires_grad_(True)
        loss, ld, lc = O.iqsl_loss(pred, torch.from_numpy(z[f"iq_tgt{i}"]), t1, t2, tau, margin, cef)
        loss.backward()
        assert np.allclose([loss.item(), ld.item(), lc.item()], z[f"iq_loss{i}"], rtol=1e-6)
        assert np.allclose(pred.grad.numpy(), z[f"iq_grad{i}"], atol=1e-9, rtol=1e-5)


def test_resnet_forward_and_live_step_match_reference(golden):
    """arch_unet.RESNET (arch_unet.py:263-409) + the fork's live supervised step (train.py:361-368)."""
    z = golden("r2_resnet")
    for tag, (in_nc, nf) in {"g1": (1, 4), "c3": (3, 8)}.items():
        seed = int(z[f"{tag}_seed"])
        p = _rand_bias(O.resnet_init(in_nc, in_nc, nf, seed), seed + 100)
        noisy = torch.from_numpy(z[f"{tag}_noisy"]); clean = torch.from_numpy(z[f"{tag}_clean"])
        with torch.no_grad():
            assert np.abs(O.resnet_forward(p, noisy).numpy() - z[f"{tag}_y"]).max() < 1e-6
        pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        loss, _, _, _ = O.structure_loss(O.resnet_forward(pr, noisy), O.resnet_forward(pr, clean), clean)
        loss.backward()
        assert abs(loss.item() - float(z[f"{tag}_loss"])) < 1e-6
        for k, v in pr.items():
            assert bool(z[f"{tag}_hasgrad/{k}"]) == (v.grad is not None), k          # up5 is constructed but unused
            if v.grad is not None:
                assert np.allclose(_csum(v.grad), z[f"{tag}_gsum/{k}"], rtol=1e-4, atol=1e-9), k
    z = golden("r2_live_step")
    p = _rand_bias(O.unet_init(1, 1, 4, 3), 103)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    noisy = torch.from_numpy(z["noisy"]); clean = torch.from_numpy(z["clean"])
    out = O.unet_forward(pr, noisy)
    loss, px, _, _ = O.structure_loss(out, O.unet_forward(pr, clean), clean)
    assert np.allclose([loss.item(), px.item()], z["losses"][0], rtol=1e-6)


def test_improved_unet_oracle_matches_reference(golden):
    """arch_unet.ImprovedUNet (arch_unet.py:420-531): forward, live-step loss and gradient checksums vs the unmodified reference;
    the drop-in module registers the reference's parameters (names, shapes, order)."""
    from image_denoising_b200 import ImprovedUNet
    z = golden("r2_improved")
    for tag in ("g16", "c48"):
        in_nc, nf, seed = (int(v) for v in z[f"{tag}_cfg"])
        p = O.improved_init(in_nc, in_nc, nf, seed)
        net = ImprovedUNet(in_nc=in_nc, out_nc=in_nc, n_feature=nf)
        assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(v.shape)) for k, v in p.items()]
        noisy = torch.from_numpy(z[f"{tag}_noisy"]); clean = torch.from_numpy(z[f"{tag}_clean"])
        with torch.no_grad():
            assert np.abs(O.improved_forward(p, noisy).numpy() - z[f"{tag}_y"]).max() < 1e-6
        if tag == "g16":
            pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
            loss, _, _, _ = O.structure_loss(O.improved_forward(pr, noisy), O.improved_forward(pr, clean), clean)
            loss.backward()
            assert abs(loss.item() - float(z[f"{tag}_loss"])) < 1e-6
            for k, v in pr.items():
                assert np.allclose(_csum(v.grad), z[f"{tag}_gsum/{k}"], rtol=1e-3, atol=1e-8), k
    assert sum(v.numel() for v in ImprovedUNet(1, 1, 48).state_dict().values()) == sum(
        int(np.prod(s)) for s in O.improved_param_shapes(1, 1, 48).values())
