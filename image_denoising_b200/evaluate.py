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
                  images_per_batch: int = 8, return_device: bool = False) -> Tuple[List[np.ndarray], List[float]]:
    """evaluation_704.py:70-120 for a list of equally-sized 2-D uint8 images.  The images are uploaded as uint8; tiling
    (cut, /255, numpy-style reflect padding), the forward of all tiles of ``images_per_batch`` images as ONE batch, the
    triangular blend in the reference's tile order and the truncating quantiser all run on the device."""
    stride = ps - overlap
    wm = torch.from_numpy(tile_weight(ps)).to(device)
    outs, l1s = [], []
    h, w = (int(v) for v in noisy_imgs[0].shape)
    nt = len(tile_origins(h, w, ps, overlap))
    for b0 in range(0, len(noisy_imgs), images_per_batch):
        chunk = noisy_imgs[b0:b0 + images_per_batch]
        if isinstance(chunk[0], torch.Tensor):
            imgs = torch.stack([c.to(device=device, dtype=torch.uint8) for c in chunk])
        else:
            imgs = torch.from_numpy(np.stack([np.asarray(n).astype(np.uint8) for n in chunk])).to(device)
        x = ops.tile_gather_u8(imgs, ps, stride)                      # [B*nt, 1, ps, ps]
        pred = network(x)
        out = ops.tile_blend_u8(pred, wm, imgs.shape[0], h, w, ps, stride)
        outs.append(out)
        for i in range(imgs.shape[0]):                                # L1(pred tile, noisy tile), evaluation_704.py:99-101
            loss3, _ = ops.l1grad_loss_fwdbwd(pred[i * nt:(i + 1) * nt], x[i * nt:(i + 1) * nt], 0.0, 1.0, want_grad=False)
            l1s.append(loss3)
    l1s = [float(t[1]) for t in torch.stack(l1s).cpu()]
    if return_device:
        return [o for batch in outs for o in batch], l1s
    return [o for batch in outs for o in batch.cpu().numpy()], l1s
