"""§8f rows N1 / N3 on the GPU: the fork's live supervised training step (train.py:354-368: network(noisy) and
network(clean) with grad, util.Structure_loss, Adam) and arch_unet.RESNET, through the drop-in modules, against golden
vectors produced by the unmodified reference (oracle/make_golden_r2.py) and the CPU oracle."""
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


def _rand_bias(p, seed):
    g = torch.Generator().manual_seed(seed)
    for k in p:
        if k.endswith(".bias"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.05
    return p


def _csum(t):
    t = t.detach().double().cpu()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def test_structure_loss_kernel_matches_reference_golden(dev, golden):
    from image_denoising_b200 import Structure_loss
    z = golden("r2_misc")
    crit = Structure_loss()
    for i in range(3):
        pred = torch.from_numpy(z[f"sl_pred{i}"]).to(dev).requires_grad_(True)
        pred2 = torch.from_numpy(z[f"sl_pred2{i}"]).to(dev).requires_grad_(True)
        tgt = torch.from_numpy(z[f"sl_tgt{i}"]).to(dev)
        loss = crit(pred, pred2, tgt)
        (3.0 * loss).backward()
        assert np.allclose(crit.last_terms.cpu().numpy(), z[f"sl_loss{i}"], rtol=2e-6)
        assert np.allclose(pred.grad.cpu().numpy(), 3.0 * z[f"sl_g1_{i}"], rtol=1e-5, atol=1e-7)
        assert np.allclose(pred2.grad.cpu().numpy(), 3.0 * z[f"sl_g2_{i}"], rtol=1e-5, atol=1e-7)
    with pytest.raises(NotImplementedError):
        Structure_loss(reduction='sum')


def test_structure_loss_large_vs_oracle(dev):
    from image_denoising_b200 import ops
    g = torch.Generator().manual_seed(2)
    a, b, t = (torch.rand(4, 1, 256, 256, generator=g) for _ in range(3))
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    lo, px, tv, cs = O.structure_loss(ar, br, t, 1.0, 0.5, 0.5)
    lo.backward()
    loss4, g1, g2 = ops.structure_loss_fwdbwd(a.to(dev), b.to(dev), t.to(dev), 1.0, 0.5, 0.5)
    assert np.allclose(loss4.cpu().numpy(), [lo.item(), px.item(), tv.item(), cs.item()], rtol=1e-5)
    assert torch.allclose(g1.cpu(), ar.grad, atol=1e-9, rtol=1e-5) and torch.allclose(g2.cpu(), br.grad, atol=1e-9, rtol=1e-5)


def test_iqsl_loss_kernel_matches_reference_golden(dev, golden):
    """finetune_iqsl.py:291-383: loss terms and dL/dpred of the fused kernels vs the reference's own function."""
    from image_denoising_b200 import iqsl_loss
    z = golden("r2_misc")
    for i in range(3):
        t1, t2, tau, margin, cef = (float(v) for v in z[f"iq_cfg{i}"])
        pred = torch.from_numpy(z[f"iq_pred{i}"]).to(dev).requires_grad_(True)
        tgt = torch.from_numpy(z[f"iq_tgt{i}"]).to(dev)
        loss = iqsl_loss(pred, tgt, t1=t1, t2=t2, tau=tau, margin=margin, ce_factor=cef)
        (2.0 * loss).backward()
        assert abs(loss.item() - z[f"iq_loss{i}"][0]) <= 2e-6 * abs(z[f"iq_loss{i}"][0])
        assert np.allclose(pred.grad.cpu().numpy(), 2.0 * z[f"iq_grad{i}"], rtol=2e-4, atol=2e-8)
    # [B,H,W] inputs are accepted like the reference; multi-channel input and mismatched shapes are rejected
    p3 = torch.rand(2, 16, 16, device=dev)
    assert iqsl_loss(p3, p3.clone(), 0.3, 0.7).dim() == 0
    with pytest.raises(ValueError):
        iqsl_loss(torch.rand(1, 3, 8, 8, device=dev), torch.rand(1, 3, 8, 8, device=dev), 0.3, 0.7)
    with pytest.raises(ValueError):
        iqsl_loss(torch.rand(1, 1, 8, 8, device=dev), torch.rand(1, 1, 8, 4, device=dev), 0.3, 0.7)


def test_iqsl_loss_large_vs_oracle(dev):
    """Finetune shape (32 x 1 x 256 x 256): the nine global Dice sums couple 2 M pixels; fp64 two-stage reduction."""
    from image_denoising_b200 import ops
    g = torch.Generator().manual_seed(5)
    a, t = torch.rand(32, 1, 256, 256, generator=g), torch.rand(32, 1, 256, 256, generator=g)
    for margin, cef in ((0.0, 0.5), (0.02, 1.0)):
        ar = a.clone().double().requires_grad_(True)
        lo, ld, lc = O.iqsl_loss(ar, t.double(), 0.35, 0.66, 0.1, margin, cef)
        lo.backward()
        loss3, grad = ops.iqsl_loss_fwdbwd(a.to(dev), t.to(dev), 0.35, 0.66, 0.1, margin, cef)
        assert np.allclose(loss3.cpu().numpy(), [lo.item(), ld.item(), lc.item()], rtol=2e-5)
        gr = ar.grad.float()
        assert (grad.cpu() - gr).abs().max().item() <= 2e-4 * gr.abs().max().item()
        loss3b, none = ops.iqsl_loss_fwdbwd(a.to(dev), t.to(dev), 0.35, 0.66, 0.1, margin, cef, want_grad=False)
        assert none is None and torch.equal(loss3b, loss3)          # deterministic, workspace counter reset


def test_space_to_depth_matches_reference_golden(dev, golden):
    from image_denoising_b200 import space_to_depth
    z = golden("r2_misc")
    for i in range(4):
        y = space_to_depth(torch.from_numpy(z[f"s2d_x{i}"]).to(dev), int(z[f"s2d_bs{i}"]))
        assert np.array_equal(y.cpu().numpy(), z[f"s2d_y{i}"])


def test_unet_live_supervised_step_fp32_matches_reference_golden(dev, golden):
    """train.py:361-368 on UNet: two forwards with grad before one backward, Structure_loss, Adam — 3 iterations."""
    from image_denoising_b200 import FusedAdam, Structure_loss, UNet
    z = golden("r2_live_step")
    net = UNet(in_nc=1, out_nc=1, n_feature=4)
    net.load_state_dict(_rand_bias(O.unet_init(1, 1, 4, 3), 103))
    net = net.to(dev).set_precision("fp32")
    noisy = torch.from_numpy(z["noisy"]).to(dev); clean = torch.from_numpy(z["clean"]).to(dev)
    opt = FusedAdam(net.parameters(), lr=3e-4)
    crit = Structure_loss()
    losses = []
    for it in range(3):
        opt.zero_grad()
        noisy_output, clean_ = net(noisy), net(clean)
        loss = crit(noisy_output, clean_, clean)
        loss.backward()
        if it == 0:
            for k, v in net.named_parameters():
                assert np.allclose(_csum(v.grad), z["gsum/" + k], rtol=2e-4, atol=1e-9), k
                if ("grad/" + k) in z.files:
                    ref = z["grad/" + k]
                    assert np.abs(v.grad.cpu().numpy() - ref).max() <= 1e-7 + 3e-4 * np.abs(ref).max(), k
        opt.step()
        losses.append([float(crit.last_terms[0]), float(crit.last_terms[1])])
    assert np.allclose(losses, z["losses"], rtol=2e-5)
    for k, v in net.state_dict().items():
        assert np.allclose(_csum(v), z["w3sum/" + k], rtol=1e-4, atol=1e-8), k


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,in_nc,nf", [("g1", 1, 4), ("c3", 3, 8)])
def test_resnet_forward_and_live_step_match_reference_golden(dev, golden, tag, in_nc, nf, precision):
    from image_denoising_b200 import RESNET, Structure_loss
    z = golden("r2_resnet")
    seed = int(z[f"{tag}_seed"])
    p = _rand_bias(O.resnet_init(in_nc, in_nc, nf, seed), seed + 100)
    net = RESNET(in_nc=in_nc, out_nc=in_nc, n_feature=nf)
    assert list(net.state_dict().keys()) == list(p.keys())
    net.load_state_dict(p)
    net = net.to(dev).set_precision(precision)
    noisy = torch.from_numpy(z[f"{tag}_noisy"]).to(dev); clean = torch.from_numpy(z[f"{tag}_clean"]).to(dev)
    with torch.no_grad():
        y = net(noisy)
    err = np.abs(y.cpu().numpy() - z[f"{tag}_y"]).max()
    assert err < (3e-6 if precision == "fp32" else 3e-2), err
    crit = Structure_loss()
    loss = crit(net(noisy), net(clean), clean)
    loss.backward()
    assert abs(float(loss) - float(z[f"{tag}_loss"])) <= (2e-5 if precision == "fp32" else 2e-2) * float(z[f"{tag}_loss"])
    for k, v in net.named_parameters():
        if not bool(z[f"{tag}_hasgrad/{k}"]):
            assert v.grad is None, k                       # up5 is registered but unused (arch_unet.py:303)
            continue
        if ("%s_grad/%s" % (tag, k)) in z.files:
            ref = z[f"{tag}_grad/{k}"].astype(np.float64); got = v.grad.cpu().numpy().astype(np.float64)
            if precision == "fp32":
                assert np.abs(got - ref).max() <= 1e-8 + 3e-4 * np.abs(ref).max(), k
            else:
                cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-300))
                assert cos > 0.97, (k, cos)
        if precision == "fp32":
            assert np.allclose(_csum(v.grad), z[f"{tag}_gsum/{k}"], rtol=5e-4, atol=1e-9), k


def test_resnet_nf48_bf16_vs_oracle(dev):
    from image_denoising_b200 import RESNET
    p = _rand_bias(O.resnet_init(1, 1, 48, 9), 10)
    g = torch.Generator().manual_seed(4)
    clean = torch.rand(2, 1, 64, 96, generator=g)
    x = clean + torch.randn(clean.shape, generator=g) * (25 / 255)
    net = RESNET(1, 1, 48)
    net.load_state_dict(p)
    with torch.no_grad():
        ref = O.resnet_forward(p, x)
        y32 = net.to(dev).set_precision("fp32")(x.to(dev)).cpu()
        y16 = net.set_precision("bf16")(x.to(dev)).cpu()
    assert float((y32 - ref).abs().max()) < 2e-5
    psnr = lambda a: 10 * np.log10(1.0 / float(((a - clean) ** 2).mean()))
    assert abs(psnr(y16) - psnr(ref)) < 0.01


@pytest.mark.parametrize("family", ["UNet", "RESNET", "ImprovedUNet"])
def test_forward_pair_equals_two_calls(dev, family):
    """entry/train.py runs train.py:361's network(noisy), network(clean) as one pass over the concatenated batch: outputs and
    parameter gradients must be those of the two calls (every layer is per-sample)."""
    import image_denoising_b200 as M
    torch.manual_seed(1)
    net = getattr(M, family)(1, 1, 16).to(dev).set_precision("fp32")
    g = torch.Generator().manual_seed(2)
    a = torch.rand(2, 1, 64, 64, generator=g).to(dev); b = torch.rand(2, 1, 64, 64, generator=g).to(dev)
    crit = M.Structure_loss()
    ya, yb = net(a), net(b)
    crit(ya, yb, b).backward()
    g2 = {k: v.grad.clone() for k, v in net.named_parameters() if v.grad is not None}
    net.zero_grad()
    pa, pb = M.forward_pair(net, a, b)
    assert (pa - ya).abs().max().item() < 1e-5 and (pb - yb).abs().max().item() < 1e-5
    crit(pa, pb, b).backward()
    for k, v in net.named_parameters():
        if v.grad is not None:
            assert (v.grad - g2[k]).abs().max().item() <= 2e-3 * g2[k].abs().max().item() + 1e-9, k
