"""Drop-in ``utils_eval`` (reference utils_eval.py:7-53): same function names / argument
meaning, PSNR and SSIM computed by the n2n_psnr_ssim_u8 reduction kernel on the GPU."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch
from PIL import Image

from . import ops


def validation_denoise(dataset_dir):
    """utils_eval.py:7-17."""
    clean = sorted(glob.glob(os.path.join(dataset_dir, 'clean', "*")))
    noise = sorted(glob.glob(os.path.join(dataset_dir, 'noise', "*")))
    images1, images2 = [], []
    for fn1, fn2 in zip(clean, noise):
        images1.append(np.array(Image.open(fn1), dtype=np.float32))
        images2.append(np.array(Image.open(fn2), dtype=np.float32))
    return images1, images2, clean, noise


def _as_u8_cuda(img, device):
    if isinstance(img, torch.Tensor):
        t = img
    else:
        a = np.asarray(img)
        if a.dtype != np.uint8:
            if not np.array_equal(a, np.floor(a)) or a.min() < 0 or a.max() > 255:
                raise ValueError("psnr/ssim kernels take 8-bit images (integral values in 0..255), as the "
                                 "reference's evaluation scripts pass them (evaluation.py:83-97)")
            a = a.astype(np.uint8)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype != torch.uint8:
        raise ValueError("expected uint8 image data")
    return t.to(device)


def psnr_ssim_batch(pred, ref, device="cuda"):
    """Batched form: lists/arrays of HxW or HxWxC uint8 images -> float64 array [B, 2]."""
    a = torch.stack([_as_u8_cuda(p, device) for p in pred])
    b = torch.stack([_as_u8_cuda(r, device) for r in ref])
    return ops.psnr_ssim_u8(a, b).cpu().numpy()


def calculate_ssim(target, ref):
    """utils_eval.py:35-47."""
    t = np.asarray(target); r = np.asarray(ref)
    if t.shape != r.shape:
        raise ValueError('Input images must have the same dimensions.')
    if t.ndim == 3 and t.shape[2] not in (1, 3):
        return None                       # the reference falls through and returns None here
    if t.ndim not in (2, 3):
        raise ValueError('Wrong input image dimensions.')
    return float(psnr_ssim_batch([t], [r])[0, 1])


def calculate_psnr(target, ref):
    """utils_eval.py:49-53."""
    return float(psnr_ssim_batch([np.asarray(target)], [np.asarray(ref)])[0, 0])


def compute_iq_iou(pred255, clean255, low_q: float, high_q: float):
    """evaluation_704_iqsl.py:53-83: per-class IoU (dark / mid / bright) of the 3-level intensity quantisation of ``pred255``
    against ``clean255``, thresholds = the (low_q, high_q) quantiles of the clean image.  Host numpy, as in the reference."""
    def gray01(img):
        arr = np.asarray(img).astype(np.float32)
        if arr.ndim == 3:
            arr = arr.mean(axis=2)
        return arr / 255.0
    gt, pr = gray01(clean255), gray01(pred255)
    t1, t2 = np.quantile(gt, [low_q, high_q])
    def quant(g):
        lab = np.zeros_like(g, dtype=np.int32)
        lab[g <= t1] = 0
        lab[(g > t1) & (g < t2)] = 1
        lab[g >= t2] = 2
        return lab
    gl, pl = quant(gt), quant(pr)
    ious = []
    for k in range(3):
        inter = np.logical_and(gl == k, pl == k).sum()
        union = np.logical_or(gl == k, pl == k).sum()
        ious.append(float("nan") if union == 0 else float(inter) / float(union))
    return ious
