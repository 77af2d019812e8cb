"""Fused ConvTranspose2x2 -> conv3x3 (no-grad bf16 passes: the full-resolution pass of the N2N step, evaluation,
the adapter's frozen base): composite weights + border bias correction against (a) the layer-by-layer launches it
replaces (N2N_NO_UPFUSE=1) and (b) the fp32 oracle, with NON-ZERO ConvTranspose biases so that the image-border
correction (the 3x3 conv zero-pads the UPSAMPLED tensor) is exercised, on shapes whose levels are not multiples of
the 8x16 tile."""
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


def _weights(in_nc, nf, seed, bias_scale=0.05):
    p = O.unet_init(in_nc, in_nc, nf, seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k in p:
        if k.endswith(".bias"):
            p[k] = torch.randn(p[k].shape, generator=g) * bias_scale
        else:
            p[k] = p[k] * 6.0          # larger activations deep in the net: border effects are not lost in the noise
    return p


def _net(dev, in_nc, nf, params):
    from image_denoising_b200 import UNet
    net = UNet(in_nc=in_nc, out_nc=in_nc, n_feature=nf)
    net.load_state_dict(params)
    return net.to(dev).set_precision("bf16")


@pytest.mark.parametrize("in_nc,nf,shape", [(1, 48, (2, 64, 96)), (3, 48, (1, 96, 160)), (1, 48, (1, 352, 352)), (1, 16, (3, 32, 64)),
                                            (1, 32, (2, 64, 64)), (3, 32, (1, 64, 96)), (1, 4, (2, 64, 64))])
def test_fused_upconv_matches_layerwise_and_oracle(dev, monkeypatch, in_nc, nf, shape):
    p = _weights(in_nc, nf, 11)
    n, h, w = shape
    g = torch.Generator().manual_seed(5)
    x = torch.rand(n, in_nc, h, w, generator=g)
    with torch.no_grad():
        ref = O.unet_forward(p, x)
    outs = {}
    for fused in (True, False):
        if fused:
            monkeypatch.delenv("N2N_NO_UPFUSE", raising=False)
        else:
            monkeypatch.setenv("N2N_NO_UPFUSE", "1")
        net = _net(dev, in_nc, nf, p)
        with torch.no_grad():
            outs[fused] = net(x.to(dev)).cpu()
        launches = net.last_launches
        outs[(fused, "launches")] = launches
    assert outs[(True, "launches")] < outs[(False, "launches")], "the fused path did not engage"
    scale = float(ref.abs().max())
    e_f = (outs[True] - ref).abs()
    e_u = (outs[False] - ref).abs()
    # same accuracy class as the layer-by-layer bf16 path, everywhere ...
    assert float(e_f.max()) <= max(2.0 * float(e_u.max()), 0.02 * scale), (float(e_f.max()), float(e_u.max()), scale)
    assert float(e_f.mean()) <= 1.5 * float(e_u.mean()) + 1e-6
    # ... and in particular on the image border (first / last two rows and columns), where the correction acts
    border = torch.zeros_like(ref, dtype=torch.bool)
    border[..., :2, :] = True; border[..., -2:, :] = True; border[..., :, :2] = True; border[..., :, -2:] = True
    assert float(e_f[border].mean()) <= 1.5 * float(e_u[border].mean()) + 1e-6, (float(e_f[border].mean()), float(e_u[border].mean()))
    mse = lambda a: float(((a - ref) ** 2).mean())
    assert mse(outs[True]) <= 2.0 * mse(outs[False]) + 1e-12


def test_fused_upconv_border_correction_is_needed(dev):
    """Sanity of the test itself: with a large ConvTranspose bias the border correction changes the output by far more
    than the bf16 tolerance, so a missing / wrong correction cannot pass the test above."""
    p = _weights(1, 48, 3, bias_scale=0.0)
    for k in p:
        if "deconv.bias" in k:
            p[k] = torch.full(p[k].shape, 0.5)
    x = torch.rand(1, 1, 64, 64, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = O.unet_forward(p, x)
        y = _net(dev, 1, 48, p)(x.to(dev)).cpu()
    interior_err = float((y - ref)[..., 4:-4, 4:-4].abs().mean())
    border_err = float(torch.cat([(y - ref)[..., 0, :].flatten(), (y - ref)[..., -1, :].flatten(),
                                  (y - ref)[..., :, 0].flatten(), (y - ref)[..., :, -1].flatten()]).abs().mean())
    assert border_err <= 3.0 * interior_err + 1e-4, (border_err, interior_err)


@pytest.mark.parametrize("nf", [4, 16, 32])
def test_bf16_training_step_other_widths_vs_oracle(dev, nf):
    """Widths whose decoder concat is not a whole number of 48-channel groups (n_feature 4 / 16 / 32): the im2col form of
    dec_conv1a keeps its extra weight slab behind nine partially filled tap slabs — a workspace sizing bug once let the
    next layer's pack overwrite it (garbage / NaN activations from dec_conv1a on)."""
    from image_denoising_b200 import N2NTrainer, UNet
    p = _weights(1, nf, 21, bias_scale=0.02)
    g = torch.Generator().manual_seed(9)
    clean = torch.rand(2, 1, 64, 64, generator=g)
    noisy = clean + torch.randn(clean.shape, generator=g) * (25 / 255)
    rd = O.draw_rd_idx(2, 64, 64, 1)
    m1, m2 = O.masks_from_rd_idx(rd)
    loss, _, _, grads, _, _ = O.n2n_step_grads(p, noisy, m1, m2, 1.0)
    net = UNet(1, 1, nf); net.load_state_dict(p); net = net.to(dev).set_precision("bf16")
    tr = N2NTrainer(net, lr=0.0, precision="bf16")
    loss3 = tr.step(noisy.to(dev), 1.0, rd_idx=torch.from_numpy(rd).to(dev)).cpu().numpy()
    assert np.isfinite(loss3).all() and abs(loss3[0] - loss) <= 3e-2 * abs(loss), (loss3, loss)
    for (k, ref), gv in zip(grads.items(), tr.grads):
        a, b = gv.cpu().double().flatten(), ref.double().flatten()
        assert torch.isfinite(a).all(), k
        cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-300))
        assert cos > 0.97, (k, cos)


@pytest.mark.parametrize("in_nc,nf,shape", [(1, 48, (3, 64, 96)), (1, 48, (5, 128, 128)), (3, 48, (2, 64, 64)), (1, 32, (2, 64, 64))])
def test_composite_backward_matches_layerwise_and_oracle(dev, monkeypatch, in_nc, nf, shape):
    """Training pass with the fused up-conv levels: forward on composite weights, input gradient with the transposed
    composites, composite weight gradient + chain rule back to dec_conv{k}a / up{k} (incl. the ConvTranspose bias gradient
    through border-aware sums of dL/dy) — every parameter gradient against the layer-by-layer bf16 backward
    (N2N_NO_UPFUSE_TRAIN=1) and the fp32 oracle, with non-zero biases everywhere."""
    p = _weights(in_nc, nf, 7, bias_scale=0.05)
    n, h, w = shape
    g = torch.Generator().manual_seed(3)
    x = torch.rand(n, in_nc, h, w, generator=g)
    tgt = torch.rand(n, in_nc, h, w, generator=g)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    (O.unet_forward(pr, x) - tgt).square().mean().backward()
    res = {}
    for fused in (True, False):
        if fused:
            monkeypatch.delenv("N2N_NO_UPFUSE_TRAIN", raising=False)
        else:
            monkeypatch.setenv("N2N_NO_UPFUSE_TRAIN", "1")
        net = _net(dev, in_nc, nf, p)
        xd = x.to(dev).requires_grad_(in_nc == 3)
        loss = (net(xd) - tgt.to(dev)).square().mean()
        loss.backward()
        res[fused] = ({k: v.grad.detach().cpu().double() for k, v in net.named_parameters()}, float(loss),
                      xd.grad.detach().cpu().double() if in_nc == 3 else None)
    assert abs(res[True][1] - res[False][1]) <= 2e-3 * abs(res[False][1])
    worst = 1.0
    for k in res[True][0]:
        a, b, r = res[True][0][k].flatten(), res[False][0][k].flatten(), pr[k].grad.double().flatten()
        assert torch.isfinite(a).all(), k
        cos_ab = float((a * b).sum() / (a.norm() * b.norm() + 1e-300))
        cos_ar = float((a * r).sum() / (a.norm() * r.norm() + 1e-300))
        cos_br = float((b * r).sum() / (b.norm() * r.norm() + 1e-300))
        worst = min(worst, cos_ar)
        # as close to the fp32 oracle as the layer-by-layer bf16 backward is (both carry bf16 activations / gradients)
        assert cos_ar >= min(0.98, cos_br - 0.01), (k, cos_ar, cos_br, cos_ab)
        assert abs(float(a.norm() / (r.norm() + 1e-300)) - 1.0) < 0.05, (k, float(a.norm()), float(r.norm()))
    if in_nc == 3:
        a, b = res[True][2].flatten(), res[False][2].flatten()
        assert float((a * b).sum() / (a.norm() * b.norm())) > 0.995
    print(f"composite backward: worst cosine vs fp32 oracle {worst:.5f}")
