"""CPU restatement (numpy + plain PyTorch fp32 functional ops) of the reference's
Neighbor2Neighbor hot path.  TEST INFRASTRUCTURE — never imported by the product.

Every function cites the reference file:line (paths relative to the reference
repo ``lmh9507/image_denoising``) whose arithmetic it restates.  The restatement
is pinned against the *unmodified* reference code by ``oracle/make_golden.py``
(run in the build container where ``/root/reference`` is mounted); the resulting
vectors live in ``tests/golden/`` and are re-checked by ``tests/test_oracle.py``
on every run.  The reference itself ships no tests / golden vectors,
so "pinned" here means "bit-equal (integer paths) or <=1e-6 (fp32 paths) to the
reference's own code executed on the same inputs".
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# a4 / a5: neighbour sub-sampler  (train.py:141-190)
# ----------------------------------------------------------------------------
# Cell-local positions: k = 2*ky + kx, k=0:(0,0) 1:(0,1) 2:(1,0) 3:(1,1)
# (train.py:134-138: F.unfold channel order).  The eight admissible (k1, k2)
# neighbour pairs, in the order the reference indexes them with rd_idx
# (train.py:151-154).
PAIR_TABLE = np.array(
    [[0, 1], [0, 2], [1, 3], [2, 3], [1, 0], [2, 0], [3, 1], [3, 2]], dtype=np.int64)


def masks_from_rd_idx(rd_idx: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """train.py:163-172 — scatter one True per 2x2 cell into two flat bool masks.

    ``rd_idx`` has one entry in [0,8) per cell, cells ordered (n, i, j) row-major.
    Returns two 1-D bool arrays of length 4*cells.
    """
    rd_idx = np.asarray(rd_idx, dtype=np.int64).reshape(-1)
    cells = rd_idx.shape[0]
    sel = PAIR_TABLE[rd_idx]                      # [cells, 2]
    m1 = np.zeros((cells, 4), dtype=np.bool_)
    m2 = np.zeros((cells, 4), dtype=np.bool_)
    m1[np.arange(cells), sel[:, 0]] = True
    m2[np.arange(cells), sel[:, 1]] = True
    return m1.reshape(-1), m2.reshape(-1)


def draw_rd_idx(n: int, h: int, w: int, seed: int) -> np.ndarray:
    """train.py:155-162 with the CPU generator variant of get_generator
    (training_script.md:4-10): a fresh generator seeded with the operation
    counter draws randint(0, 8) per cell."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    cells = n * h // 2 * w // 2      # same left-to-right precedence as train.py:144
    return torch.randint(0, 8, (cells,), generator=g, dtype=torch.int64).numpy()


def add_train_noise_gauss(x: torch.Tensor, sigma255: float, seed: int) -> torch.Tensor:
    """train.py:84-94 (gauss_fix) with the generator of training_script.md:4-10 seeded with the operation
    counter value ``seed``: x + N(0, (sigma/255)^2), per-sample std tensor, no clamp.
    PARITY NOTE: ``torch.normal(mean, std[B,1,1,1], out=noise[B,C,H,W])`` fills ``noise`` at its own shape under the
    reference's pinned PyTorch 1.3 (normal_(0,1) then mul_(std)); torch >= 2 resizes ``out`` to [B,1,1,1] (one offset
    per sample).  The intended per-pixel form is restated; it cannot be pinned by running the reference under torch 2.x
    ("parity unpinned" for this data-generation helper only)."""
    g = torch.Generator(device=x.device)
    g.manual_seed(int(seed))
    std = (sigma255 / 255.0) * torch.ones((x.shape[0], 1, 1, 1), device=x.device)
    noise = torch.zeros(x.shape, dtype=torch.float32, device=x.device)
    noise.normal_(0.0, 1.0, generator=g)
    return x + noise * std


def space_to_depth(x: np.ndarray, bs: int) -> np.ndarray:
    """train.py:134-138: F.unfold(x, bs, stride=bs).view(n, c*bs*bs, h//bs, w//bs);
    channel = c*bs*bs + ky*bs + kx."""
    n, c, h, w = x.shape
    t = x.reshape(n, c, h // bs, bs, w // bs, bs)           # n c i ky j kx
    return np.ascontiguousarray(t.transpose(0, 1, 3, 5, 2, 4)).reshape(n, c * bs * bs, h // bs, w // bs)


def subimage_from_mask(img: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """train.py:175-190 — out[n,c,i,j] = img[n,c,2i+k//2,2j+k%2], k = the True slot
    of cell (n,i,j) in ``mask``; the same mask serves every channel."""
    n, c, h, w = img.shape
    hh, ww = h // 2, w // 2
    m = np.asarray(mask).reshape(n, hh, ww, 4)
    if not (m.sum(axis=-1) == 1).all():
        raise ValueError("mask must select exactly one pixel per 2x2 cell")
    k = m.argmax(axis=-1)                          # [n, hh, ww]
    ky, kx = k // 2, k % 2
    ii = 2 * np.arange(hh)[None, :, None] + ky
    jj = 2 * np.arange(ww)[None, None, :] + kx
    nn = np.arange(n)[:, None, None]
    out = img[nn, :, ii, jj]                       # [n, hh, ww, c]
    return np.ascontiguousarray(np.moveaxis(out, -1, 1))


# ----------------------------------------------------------------------------
# a6 / a7 / a8: UNet  (arch_unet.py:100-260, non-blindspot branch)
# ----------------------------------------------------------------------------
def unet_param_shapes(in_nc: int, out_nc: int, nf: int) -> "OrderedDict[str, tuple]":
    """arch_unet.py:114-192 — registration order and shapes of the 50 tensors (25 layers)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k)
        s[name + ".bias"] = (co,)

    def deconv(name, ci, co):
        s[name + ".deconv.weight"] = (ci, co, 2, 2)    # ConvTranspose2d layout
        s[name + ".deconv.bias"] = (co,)

    conv("enc_conv0", nf, in_nc, 3)
    for i in range(1, 7):
        conv(f"enc_conv{i}", nf, nf, 3)
    deconv("up5", nf, nf)
    conv("dec_conv5a", 2 * nf, 2 * nf, 3)
    conv("dec_conv5b", 2 * nf, 2 * nf, 3)
    for lvl in (4, 3, 2):
        deconv(f"up{lvl}", 2 * nf, 2 * nf)
        conv(f"dec_conv{lvl}a", 2 * nf, 3 * nf, 3)
        conv(f"dec_conv{lvl}b", 2 * nf, 2 * nf, 3)
    deconv("up1", 2 * nf, 2 * nf)
    conv("dec_conv1a", 96, 2 * nf + in_nc, 3)
    conv("dec_conv1b", 96, 96, 3)
    conv("nin_a", 96, 96, 1)
    conv("nin_b", 96, 96, 1)
    conv("nin_c", out_nc, 96, 1)
    return s


def unet_init(in_nc: int, out_nc: int, nf: int, seed: int) -> "OrderedDict[str, torch.Tensor]":
    """arch_unet.py:24-33 — kaiming_normal(fan_in, a=0) * 0.1, zero bias.  (Not
    RNG-stream-identical to the reference constructor; weights for parity tests
    are always passed explicitly.)"""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shp in unet_param_shapes(in_nc, out_nc, nf).items():
        if name.endswith(".bias"):
            p[name] = torch.zeros(shp)
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            std = math.sqrt(2.0 / fan_in)
            p[name] = torch.randn(shp, generator=g) * std * 0.1
    return p


def unet_forward(p: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """arch_unet.py:194-260 (blindspot=False)."""
    act = lambda t: F.leaky_relu(t, 0.2)
    c3 = lambda t, n: F.conv2d(t, p[n + ".weight"], p[n + ".bias"], padding=1)
    c1 = lambda t, n: F.conv2d(t, p[n + ".weight"], p[n + ".bias"])
    up = lambda t, skip, n: torch.cat(
        [F.conv_transpose2d(t, p[n + ".deconv.weight"], p[n + ".deconv.bias"], stride=2), skip], 1)
    skips = [x]
    t = act(c3(x, "enc_conv0"))
    t = F.max_pool2d(act(c3(t, "enc_conv1")), 2)
    skips.append(t)
    for i in (2, 3, 4):
        t = F.max_pool2d(act(c3(t, f"enc_conv{i}")), 2)
        skips.append(t)
    t = F.max_pool2d(act(c3(t, "enc_conv5")), 2)
    t = act(c3(t, "enc_conv6"))
    for lvl in (5, 4, 3, 2, 1):
        t = up(t, skips[lvl - 1], f"up{lvl}")
        t = act(c3(t, f"dec_conv{lvl}a"))
        t = act(c3(t, f"dec_conv{lvl}b"))
    t = act(c1(t, "nin_a"))
    t = act(c1(t, "nin_b"))
    return c1(t, "nin_c")


# ----------------------------------------------------------------------------
# a9: N2N loss and the documented training step (training_script.md:128-156)
# ----------------------------------------------------------------------------
def n2n_loss(out, sub2, den1, den2, lam: float):
    """training_script.md:146-153."""
    diff = out - sub2
    exp_diff = den1 - den2
    loss1 = torch.mean(diff ** 2)
    loss2 = lam * torch.mean((diff - exp_diff) ** 2)
    return loss1 + loss2, loss1, loss2


def n2n_step_grads(p: Dict[str, torch.Tensor], noisy: torch.Tensor,
                   mask1: np.ndarray, mask2: np.ndarray, lam: float):
    """One N2N iteration up to (and including) backward: returns
    (loss_all, loss1, loss2, grads, noisy_denoised, net_out)."""
    params = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    nz = noisy.detach().numpy()
    sub1 = torch.from_numpy(subimage_from_mask(nz, mask1))
    sub2 = torch.from_numpy(subimage_from_mask(nz, mask2))
    with torch.no_grad():
        den = unet_forward(params, noisy)
    den1 = torch.from_numpy(subimage_from_mask(den.numpy(), mask1))
    den2 = torch.from_numpy(subimage_from_mask(den.numpy(), mask2))
    out = unet_forward(params, sub1)
    loss, l1, l2 = n2n_loss(out, sub2, den1, den2, lam)
    loss.backward()
    grads = OrderedDict((k, v.grad.detach().clone()) for k, v in params.items())
    return loss.item(), l1.item(), l2.item(), grads, den, out.detach()


# ----------------------------------------------------------------------------
# a10: Adam (train.py:332 — torch.optim.Adam defaults) and MultiStepLR (:333-340)
# ----------------------------------------------------------------------------
def adam_update(p, g, m, v, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """One torch-default Adam update on float32 numpy arrays (in place).
    step is 1-based.  denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= lr/(1-b1^t) * m/denom."""
    m *= np.float32(b1); m += np.float32(1 - b1) * g
    v *= np.float32(b2); v += np.float32(1 - b2) * g * g
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = np.sqrt(v) / np.float32(math.sqrt(bc2)) + np.float32(eps)
    p -= np.float32(lr / bc1) * (m / denom)
    return p, m, v


def multistep_lr(base_lr: float, epoch: int, n_epoch: int, gamma: float) -> float:
    """LR in force during 1-based ``epoch`` (train.py:333-340, :346, :375):
    milestones int(20r)-1.. with r = n_epoch/100; scheduler.step() runs at the
    end of every epoch, so epoch e has seen e-1 scheduler steps."""
    r = n_epoch / 100
    ms = [int(20 * r) - 1, int(40 * r) - 1, int(60 * r) - 1, int(80 * r) - 1]
    k = sum(1 for m_ in ms if (epoch - 1) >= m_)
    return base_lr * gamma ** k


# ----------------------------------------------------------------------------
# a11 / a12: adapter and finetune loss (adapter.py:5-67, finetune.py:153-162,:283-285)
# ----------------------------------------------------------------------------
def adapter_forward(ap: Dict[str, torch.Tensor], noisy, base_out):
    """adapter.py:22-26."""
    t = torch.cat([noisy, base_out], 1)
    t = F.relu(F.conv2d(t, ap["adapter.net.0.weight"], ap["adapter.net.0.bias"], padding=1))
    t = F.conv2d(t, ap["adapter.net.2.weight"], ap["adapter.net.2.bias"], padding=1)
    return base_out + t


def finetune_loss(pred, clean, lambda_grad: float):
    """finetune.py:153-162, :283-285."""
    l1 = torch.mean(torch.abs(pred - clean))
    gx = torch.mean(torch.abs((pred[..., :, 1:] - pred[..., :, :-1]) - (clean[..., :, 1:] - clean[..., :, :-1])))
    gy = torch.mean(torch.abs((pred[..., 1:, :] - pred[..., :-1, :]) - (clean[..., 1:, :] - clean[..., :-1, :])))
    return l1 + lambda_grad * (gx + gy), l1, gx + gy


# ----------------------------------------------------------------------------
# a15 / a16: PSNR / SSIM  (utils_eval.py:19-53)
# ----------------------------------------------------------------------------
def gaussian_taps(ksize: int = 11, sigma: float = 1.5) -> np.ndarray:
    """cv2.getGaussianKernel(11, 1.5) (utils_eval.py:24): exp(-(i-c)^2/(2 s^2)), normalised, float64."""
    c = (ksize - 1) / 2.0
    k = np.exp(-((np.arange(ksize) - c) ** 2) / (2.0 * sigma * sigma))
    return k / k.sum()


def _valid_blur(a: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """11x11 Gaussian restricted to the interior [5:-5,5:-5] (utils_eval.py:26-31);
    the interior of filter2D does not depend on the border mode."""
    k = taps.shape[0]
    h, w = a.shape
    tmp = np.zeros((h - k + 1, w), dtype=np.float64)
    for i in range(k):
        tmp += taps[i] * a[i:i + h - k + 1, :]
    out = np.zeros((h - k + 1, w - k + 1), dtype=np.float64)
    for j in range(k):
        out += taps[j] * tmp[:, j:j + w - k + 1]
    return out


def ssim_plane(a: np.ndarray, b: np.ndarray) -> float:
    """utils_eval.py:19-33 on one 2-D plane (values on the 0..255 scale)."""
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    a = a.astype(np.float64); b = b.astype(np.float64)
    t = gaussian_taps()
    mu1, mu2 = _valid_blur(a, t), _valid_blur(b, t)
    s11 = _valid_blur(a * a, t) - mu1 * mu1
    s22 = _valid_blur(b * b, t) - mu2 * mu2
    s12 = _valid_blur(a * b, t) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2))
    return float(m.mean())


def calculate_ssim(a: np.ndarray, b: np.ndarray) -> float:
    """utils_eval.py:35-47."""
    if a.shape != b.shape:
        raise ValueError("Input images must have the same dimensions.")
    if a.ndim == 2:
        return ssim_plane(a, b)
    if a.ndim == 3 and a.shape[2] == 3:
        return float(np.mean([ssim_plane(a[:, :, i], b[:, :, i]) for i in range(3)]))
    if a.ndim == 3 and a.shape[2] == 1:
        return ssim_plane(a[:, :, 0], b[:, :, 0])
    raise ValueError("Wrong input image dimensions.")


def calculate_psnr(a: np.ndarray, b: np.ndarray) -> float:
    """utils_eval.py:49-53 — float32 arithmetic, no mse==0 guard."""
    d = a.astype(np.float32) - b.astype(np.float32)
    with np.errstate(divide="ignore"):
        return float(10.0 * np.log10(255.0 * 255.0 / np.mean(np.square(d))))


# ----------------------------------------------------------------------------
# a13 / a14: evaluation post-processing  (evaluation.py:73-83, evaluation_704.py:57-120)
# ----------------------------------------------------------------------------
def quantize_round(pred01: np.ndarray) -> np.ndarray:
    """evaluation.py:82-83: clamp(0,1) -> clip(p*255+0.5, 0, 255) -> uint8."""
    p = np.clip(pred01.astype(np.float32), 0.0, 1.0)
    return np.clip(p * np.float32(255.0) + np.float32(0.5), 0, 255).astype(np.uint8)


def tile_weight(ps: int = 352) -> np.ndarray:
    """evaluation_704.py:62-68 — separable triangular window, float32, border == 0."""
    y = np.linspace(0, 1, ps)
    w1 = 1 - np.abs(y - 0.5) * 2
    return (w1[:, None] * w1[None, :]).astype(np.float32)


def tiled_denoise(forward, noisy_u8: np.ndarray, ps: int = 352, overlap: int = 64) -> np.ndarray:
    """evaluation_704.py:74-120 for one 2-D uint8 image.  ``forward`` maps a
    float32 [1,1,ps,ps] tensor to the same shape.  Returns the uint8 result
    (truncating quantisation, no +0.5 — evaluation_704.py:120)."""
    h, w = noisy_u8.shape
    stride = ps - overlap
    wm = tile_weight(ps)
    acc = np.zeros((h, w), np.float32)
    cnt = np.zeros((h, w), np.float32)
    for r0 in range(0, h, stride):
        for c0 in range(0, w, stride):
            r1, c1 = min(r0 + ps, h), min(c0 + ps, w)
            patch = noisy_u8[r0:r1, c0:c1].astype(np.float32) / 255.0
            padded = np.pad(patch, ((0, ps - patch.shape[0]), (0, ps - patch.shape[1])), mode="reflect")
            with torch.no_grad():
                o = forward(torch.from_numpy(padded)[None, None])
            o = o[0, 0].clamp(0, 1).numpy()[:patch.shape[0], :patch.shape[1]]
            wv = wm[:patch.shape[0], :patch.shape[1]]
            acc[r0:r1, c0:c1] += o * wv
            cnt[r0:r1, c0:c1] += wv
    cnt[cnt == 0] = 1
    return np.clip(acc / cnt * 255.0, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------
# N1: Structure_loss (util.py:41-70) — the criterion of the fork's live loop (train.py:322, :361-363)
# ----------------------------------------------------------------------------
def structure_loss(pred, pred2, target, alpha: float = 1.0, beta: float = 0.5, gamma: float = 0.5):
    """util.py:56-70: alpha*L1(pred,target) + beta*(L1(dy pred2)+L1(dx pred2))/2 + gamma*L1(pred2,target).
    Returns (loss, pixel, TV, consistency)."""
    pixel = torch.mean(torch.abs(pred - target))
    tv1 = torch.mean(torch.abs(pred2[:, :, 1:, :] - pred2[:, :, :-1, :]))
    tv2 = torch.mean(torch.abs(pred2[:, :, :, 1:] - pred2[:, :, :, :-1]))
    tv = (tv1 + tv2) / 2
    cst = torch.mean(torch.abs(pred2 - target))
    return alpha * pixel + beta * tv + gamma * cst, pixel, tv, cst


# ----------------------------------------------------------------------------
# N3: RESNET (arch_unet.py:263-409, non-blindspot): the UNet's convolutions without pooling / up-sampling, all at
# full resolution, global residual (:409).  up5 is constructed (:303) but never used by forward.
# ----------------------------------------------------------------------------
def resnet_param_shapes(in_nc: int, out_nc: int, nf: int) -> "OrderedDict[str, tuple]":
    """arch_unet.py:279-347 — registration order of the 42 tensors (21 layers)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k)
        s[name + ".bias"] = (co,)

    conv("enc_conv0", nf, in_nc, 3)
    for i in range(1, 7):
        conv(f"enc_conv{i}", nf, nf, 3)
    s["up5.deconv.weight"] = (nf, nf, 2, 2)
    s["up5.deconv.bias"] = (nf,)
    conv("dec_conv5a", 2 * nf, 2 * nf, 3)
    conv("dec_conv5b", 2 * nf, 2 * nf, 3)
    for lvl in (4, 3, 2):
        conv(f"dec_conv{lvl}a", 2 * nf, 3 * nf, 3)
        conv(f"dec_conv{lvl}b", 2 * nf, 2 * nf, 3)
    conv("dec_conv1a", 96, 2 * nf + in_nc, 3)
    conv("dec_conv1b", 96, 96, 3)
    conv("nin_a", 96, 96, 1)
    conv("nin_b", 96, 96, 1)
    conv("nin_c", out_nc, 96, 1)
    return s


def resnet_init(in_nc: int, out_nc: int, nf: int, seed: int) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shp in resnet_param_shapes(in_nc, out_nc, nf).items():
        if name.endswith(".bias"):
            p[name] = torch.zeros(shp)
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            p[name] = torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_in) * 0.1
    return p


def resnet_forward(p: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """arch_unet.py:349-409 (blindspot=False)."""
    act = lambda t: F.leaky_relu(t, 0.2)
    c3 = lambda t, n: F.conv2d(t, p[n + ".weight"], p[n + ".bias"], padding=1)
    c1 = lambda t, n: F.conv2d(t, p[n + ".weight"], p[n + ".bias"])
    pool0 = x
    t = act(c3(x, "enc_conv0"))
    t = act(c3(t, "enc_conv1")); pool1 = t
    t = act(c3(t, "enc_conv2")); pool2 = t
    t = act(c3(t, "enc_conv3")); pool3 = t
    t = act(c3(t, "enc_conv4")); pool4 = t
    t = act(c3(t, "enc_conv5"))
    t = act(c3(t, "enc_conv6"))
    for lvl, skip in ((5, pool4), (4, pool3), (3, pool2), (2, pool1), (1, pool0)):
        t = torch.cat([t, skip], 1)
        t = act(c3(t, f"dec_conv{lvl}a"))
        t = act(c3(t, f"dec_conv{lvl}b"))
    t = act(c1(t, "nin_a"))
    t = act(c1(t, "nin_b"))
    return c1(t, "nin_c") + x


# ----------------------------------------------------------------------------
# N4: IQSL — intensity-quantised structural loss (finetune_iqsl.py:291-383)
# ----------------------------------------------------------------------------
def iqsl_loss(pred, target, t1: float, t2: float, tau: float = 0.1, margin: float = 0.0, ce_factor: float = 0.5,
              eps: float = 1e-6):
    """finetune_iqsl.py:291-383 on [B,1,H,W] tensors.  Returns (total, dice term, CE term)."""
    y_s, yh = target[:, 0], pred[:, 0]
    if margin > 0.0:
        valid = ((y_s <= (t1 - margin)) | ((y_s >= (t1 + margin)) & (y_s <= (t2 - margin))) | (y_s >= (t2 + margin))).float()
    else:
        valid = torch.ones_like(y_s)
    oh = torch.stack([(y_s <= t1).float(), ((y_s > t1) & (y_s < t2)).float(), (y_s >= t2).float()], dim=1)
    centers = torch.tensor([t1 / 2.0, (t1 + t2) / 2.0, (t2 + 1.0) / 2.0], dtype=pred.dtype).view(1, 3, 1, 1)
    prob = torch.softmax(-torch.abs(yh.unsqueeze(1) - centers) / max(float(tau), 1e-6), dim=1)
    vb = valid.unsqueeze(1)
    prob = prob * vb
    oh = oh * vb
    inter = (prob * oh).sum(dim=(0, 2, 3)); ps = prob.sum(dim=(0, 2, 3)); ts = oh.sum(dim=(0, 2, 3))
    loss_dice = 1.0 - ((2.0 * inter + eps) / (ps + ts + eps)).mean()
    ce = -(oh * torch.log(prob + eps)).sum() / (vb.sum() * 3 + eps)
    return loss_dice + ce_factor * ce, loss_dice, ce


# ----------------------------------------------------------------------------
# N2: ImprovedUNet (arch_unet.py:420-531)
# ----------------------------------------------------------------------------
def gn_groups(channels: int, groups: int = 32) -> int:
    """arch_unet.py:11-15 (norm2d 'gn')."""
    g = min(groups, channels)
    while channels % g != 0 and g > 1:
        g -= 1
    return g


def improved_param_shapes(in_nc: int, out_nc: int, nf0: int, depth: int = 4, noise: bool = True) -> "OrderedDict[str, tuple]":
    """state_dict keys / shapes of arch_unet.ImprovedUNet in registration order (arch_unet.py:476-513)."""
    s = OrderedDict()

    def conv(name, co, ci, k, bias=True):
        s[name + ".weight"] = (co, ci, k, k)
        if bias:
            s[name + ".bias"] = (co,)

    def rdb(pre, c):
        ci = c
        for i in range(4):
            conv(f"{pre}.convs.{i}", 32, ci, 3)
            ci += 32
        conv(f"{pre}.lff", c, ci, 1)

    def res(pre, c):
        for j in (0, 3):
            conv(f"{pre}.block.{j}", c, c, 3, bias=False)
            s[f"{pre}.block.{j + 1}.weight"] = (c,)
            s[f"{pre}.block.{j + 1}.bias"] = (c,)

    if noise:
        conv("noise_estimator.0", nf0, in_nc, 3)
        conv("noise_estimator.2", 1, nf0, 3)
    nf = nf0
    for i in range(depth):
        inc = (in_nc + 1 if noise else 1) if i == 0 else nf // 2
        conv(f"downs.{i}.0", nf, inc, 3)
        rdb(f"downs.{i}.2", nf)
        res(f"downs.{i}.3", nf)
        nf *= 2
    rdb("bottle.0", nf // 2)
    res("bottle.1", nf // 2)
    nf //= 2
    for i in range(depth):
        conv(f"ups.{i}.conv_ps", 2 * nf, nf, 3)          # out_ch * 4 = (nf // 2) * 4
        conv(f"ups.{i}.fuse", nf // 2, 3 * (nf // 2), 3)
        rdb(f"ups.{i}.rdb", nf // 2)
        res(f"ups.{i}.res", nf // 2)
        nf //= 2
    conv("final", out_nc, nf0 // 2 + in_nc, 3)
    return s


def improved_init(in_nc: int, out_nc: int, nf: int, seed: int, depth: int = 4, noise: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """Seeded test weights (NOT the reference's default init): convs ~ N(0, 1/fan_in), biases ~ 0.05 N, GroupNorm gain ~ 1 + 0.1 N."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shp in improved_param_shapes(in_nc, out_nc, nf, depth, noise).items():
        if len(shp) == 4:
            p[name] = torch.randn(shp, generator=g) * math.sqrt(1.0 / (shp[1] * shp[2] * shp[3]))
        elif name.endswith(".weight"):
            p[name] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            p[name] = 0.05 * torch.randn(shp, generator=g)
    return p


def improved_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, depth: int = 4, noise: bool = True) -> torch.Tensor:
    """arch_unet.py:515-531."""
    act = lambda t: F.leaky_relu(t, 0.2)
    in_nc = x.shape[1]

    def conv(t, n):
        w = p[n + ".weight"]
        return F.conv2d(t, w, p.get(n + ".bias"), padding=w.shape[2] // 2)

    def gn(t, n):
        return F.group_norm(t, gn_groups(t.shape[1]), p[n + ".weight"], p[n + ".bias"], 1e-5)

    def rdb(t, pre):
        feats = [t]
        for i in range(4):
            feats.append(act(conv(torch.cat(feats, 1), f"{pre}.convs.{i}")))
        return t + conv(torch.cat(feats, 1), f"{pre}.lff")

    def res(t, pre):
        u = act(gn(conv(t, f"{pre}.block.0"), f"{pre}.block.1"))
        return t + gn(conv(u, f"{pre}.block.3"), f"{pre}.block.4")

    if noise:
        sigma = torch.sigmoid(conv(act(conv(x, "noise_estimator.0")), "noise_estimator.2"))
        x = torch.cat([x, sigma], 1)
    orig = x[:, :in_nc]
    skips = []
    for i in range(depth):
        x = res(rdb(act(conv(x, f"downs.{i}.0")), f"downs.{i}.2"), f"downs.{i}.3")
        skips.append(x)
        x = F.max_pool2d(x, 2)
    x = res(rdb(x, "bottle.0"), "bottle.1")
    for i, skip in enumerate(reversed(skips)):
        x = F.pixel_shuffle(conv(x, f"ups.{i}.conv_ps"), 2)
        x = act(conv(torch.cat([x, skip], 1), f"ups.{i}.fuse"))
        x = res(rdb(x, f"ups.{i}.rdb"), f"ups.{i}.res")
    return torch.sigmoid(conv(torch.cat([x, orig], 1), "final"))
