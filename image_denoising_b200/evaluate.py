"""Evaluation paths of the reference as library functions.

* ``denoise_whole``  — evaluation.py:62-101: one full-image forward per image, clamp(0,1),
  ``clip(p*255+0.5)`` -> uint8.
* ``denoise_tiled``  — evaluation_704.py:57-120: 352x352 tiles at stride 288, edge tiles
  reflect-padded, triangular blend whose border weight is exactly 0, ``clip(p*255)`` -> uint8
  (truncation).  The nine tiles of a 704x704 image are independent, so they run as ONE batched
  forward; the blend is accumulated tile by tile in the reference's order (bit-exact fp32 sums).
* PSNR / SSIM come from the n2n_psnr_ssim_u8 reduction kernel (utils_eval.py:19-53).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import ops


def tile_weight(ps: int = 352) -> np.ndarray:
    """evaluation_704.py:62-68."""
    yy, xx = np.meshgrid(np.linspace(0, 1, ps), np.linspace(0, 1, ps), indexing="ij")
    return ((1 - np.abs(yy - 0.5) * 2) * (1 - np.abs(xx - 0.5) * 2)).astype(np.float32)


def _to_nchw(img: np.ndarray) -> np.ndarray:
    a = np.asarray(img, dtype=np.float32) / 255.0          # evaluation.py:69-70
    if a.ndim == 2:
        a = a[None]
    else:
        a = np.transpose(a, (2, 0, 1))                      # ToTensor on float HWC arrays
    return a


@torch.no_grad()
def denoise_whole(network, noisy_imgs: Sequence[np.ndarray], device="cuda") -> Tuple[List[np.ndarray], List[float]]:
    """evaluation.py:62-83 for a list of equally-sized images, batched into one forward.
    Returns (uint8 predictions in HxW / HxWxC form, per-image L1(pred, noisy_input))."""
    x = torch.from_numpy(np.stack([_to_nchw(n) for n in noisy_imgs])).to(device)
    pred = network(x)
    l1 = []
    for i in range(x.shape[0]):
        loss3, _ = ops.l1grad_loss_fwdbwd(pred[i:i + 1], x[i:i + 1], 0.0, 1.0, want_grad=False)
        l1.append(loss3)
    q = ops.quantize_u8(pred, 0.5)                          # clamp(0,1) -> clip(p*255+0.5) -> uint8
    q = q.permute(0, 2, 3, 1).cpu().numpy()
    l1 = [float(t[1]) for t in torch.stack(l1).cpu()]
    return [np.squeeze(p) for p in q], l1


def tile_origins(h: int, w: int, ps: int = 352, overlap: int = 64):
    stride = ps - overlap
    return [(r, c) for r in range(0, h, stride) for c in range(0, w, stride)]


@torch.no_grad()
def denoise_tiled(network, noisy_imgs: Sequence[np.ndarray], ps: int = 352, overlap: int = 64, device="cuda",
                  images_per_batch: int = 8) -> Tuple[List[np.ndarray], List[float]]:
    """evaluation_704.py:70-120 for a list of equally-sized 2-D uint8 images."""
    wm_host = tile_weight(ps)
    wm = torch.from_numpy(wm_host).to(device)
    outs, l1s = [], []
    h, w = np.asarray(noisy_imgs[0]).shape
    origins = tile_origins(h, w, ps, overlap)
    for b0 in range(0, len(noisy_imgs), images_per_batch):
        chunk = [np.asarray(n).astype(np.uint8) for n in noisy_imgs[b0:b0 + images_per_batch]]
        tiles, geo = [], []
        for noisy in chunk:
            for (r0, c0) in origins:
                r1, c1 = min(r0 + ps, h), min(c0 + ps, w)
                patch = noisy[r0:r1, c0:c1].astype(np.float32) / 255.0
                tiles.append(np.pad(patch, ((0, ps - patch.shape[0]), (0, ps - patch.shape[1])), mode='reflect'))
                geo.append((r0, c0, r1 - r0, c1 - c0))
        x = torch.from_numpy(np.stack(tiles)[:, None]).to(device)
        pred = network(x)
        nt = len(origins)
        for i in range(len(chunk)):
            acc = torch.zeros((h, w), dtype=torch.float32, device=device)
            cnt = torch.zeros((h, w), dtype=torch.float32, device=device)
            for t in range(nt):
                r0, c0, th, tw = geo[i * nt + t]
                ops.tile_accumulate(pred[i * nt + t, 0], wm, acc, cnt, r0, c0, th, tw)
            outs.append(ops.tile_finalize_u8(acc, cnt))
            loss3, _ = ops.l1grad_loss_fwdbwd(pred[i * nt:(i + 1) * nt], x[i * nt:(i + 1) * nt], 0.0, 1.0, want_grad=False)
            l1s.append(loss3)
    outs = [o.cpu().numpy() for o in outs]
    l1s = [float(t[1]) for t in torch.stack(l1s).cpu()]
    return outs, l1s
