"""Neighbor2Neighbor step ingredients with the reference's signatures
(train.py:56-190, training_script.md:4-10, :128-156), backed by libn2n_b200 kernels.

``generate_mask_pair(img) -> (mask1, mask2)`` and ``generate_subimages(img, mask)``
keep the reference call surface (1-D bool masks of length n*h//2*w//2*4, cell order
(n,i,j)); the random draw uses the same ``torch.randint(0, 8, generator=get_generator())``
call on the image's device, so for the same ``operation_seed_counter`` the masks are
the ones the reference would produce on that device.
"""
from __future__ import annotations

import datetime
import os

import numpy as np
import torch

from . import ops

operation_seed_counter = 0


def get_generator(device="cuda"):
    """training_script.md:4-10 (the train.py:56-61 copy lost its global counter)."""
    global operation_seed_counter
    operation_seed_counter += 1
    g = torch.Generator(device=device)
    g.manual_seed(operation_seed_counter)
    return g


class AugmentNoise(object):
    """train.py:64-131 — synthetic noise for training (torch RNG; data generation only)."""

    def __init__(self, style):
        if style.startswith('gauss'):
            self.params = [float(p) / 255.0 for p in style.replace('gauss', '', 1).split('_')]
            self.style = "gauss_fix" if len(self.params) == 1 else "gauss_range"
        elif style.startswith('poisson'):
            self.params = [float(p) for p in style.replace('poisson', '', 1).split('_')]
            self.style = "poisson_fix" if len(self.params) == 1 else "poisson_range"
        else:
            raise ValueError(f"unknown noise style {style!r}")

    def add_train_noise(self, x):
        shape = x.shape
        dev = x.device
        if self.style == "gauss_fix":
            std = self.params[0] * torch.ones((shape[0], 1, 1, 1), device=dev)
            noise = torch.empty(shape, dtype=torch.float32, device=dev)
            torch.normal(mean=0.0, std=std, generator=get_generator(dev), out=noise)
            return x + noise
        if self.style == "gauss_range":
            min_std, max_std = self.params
            std = torch.rand(size=(shape[0], 1, 1, 1), device=dev) * (max_std - min_std) + min_std
            noise = torch.empty(shape, dtype=torch.float32, device=dev)
            torch.normal(mean=0, std=std, generator=get_generator(dev), out=noise)
            return x + noise
        if self.style == "poisson_fix":
            lam = self.params[0] * torch.ones((shape[0], 1, 1, 1), device=dev)
            return torch.poisson(lam * x, generator=get_generator(dev)) / lam
        min_lam, max_lam = self.params
        lam = torch.rand(size=(shape[0], 1, 1, 1), device=dev) * (max_lam - min_lam) + min_lam
        return torch.poisson(lam * x, generator=get_generator(dev)) / lam

    def add_valid_noise(self, x):
        shape = x.shape
        if self.style == "gauss_fix":
            return np.array(x + np.random.normal(size=shape) * self.params[0], dtype=np.float32)
        if self.style == "gauss_range":
            std = np.random.uniform(low=self.params[0], high=self.params[1], size=(1, 1, 1))
            return np.array(x + np.random.normal(size=shape) * std, dtype=np.float32)
        if self.style == "poisson_fix":
            lam = self.params[0]
            return np.array(np.random.poisson(lam * x) / lam, dtype=np.float32)
        lam = np.random.uniform(low=self.params[0], high=self.params[1], size=(1, 1, 1))
        return np.array(np.random.poisson(lam * x) / lam, dtype=np.float32)


def draw_rd_idx(img: torch.Tensor) -> torch.Tensor:
    """train.py:155-162: one randint(0, 8) per 2x2 cell from a fresh counter-seeded generator."""
    n, c, h, w = img.shape
    cells = n * h // 2 * w // 2
    rd_idx = torch.zeros(size=(cells,), dtype=torch.int64, device=img.device)
    torch.randint(low=0, high=8, size=(cells,), generator=get_generator(img.device), out=rd_idx)
    return rd_idx


def generate_mask_pair(img):
    """train.py:141-172."""
    m1, m2, _ = ops.mask_pair_from_rdidx(draw_rd_idx(img), want_masks=True)
    return m1, m2


def generate_packed_selector(img):
    """Fast form of generate_mask_pair: one byte per cell (k1 | k2 << 2), same random draw."""
    _, _, pk = ops.mask_pair_from_rdidx(draw_rd_idx(img), want_masks=False, want_packed=True)
    return pk


def generate_subimages(img, mask):
    """train.py:175-190."""
    return ops.subsample(img, mask)


def generate_subimage_pair(img, mask1=None, mask2=None, packed=None):
    """Both sub-images in one pass (what the N2N step actually needs)."""
    return ops.subsample_pair(img, mask1, mask2, packed)


def space_to_depth(x, block_size):
    """train.py:134-138 — kept for API parity; only block_size == 2 is on the hot path and the
    kernels never materialise it."""
    raise NotImplementedError("space_to_depth is fused into generate_subimages; it is never materialised")


def checkpoint(net, epoch, name, save_model_path, log_name, systime=None):
    """train.py:47-53: <save_model_path>/<log_name>/<systime>/epoch_<name>_<epoch:03d>.pth."""
    systime = systime or datetime.datetime.now().strftime('%Y-%m-%d-%H-%M')
    d = os.path.join(save_model_path, log_name, systime)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, 'epoch_{}_{:03d}.pth'.format(name, epoch))
    torch.save(net.state_dict(), path)
    print('Checkpoint saved to {}'.format(path))
    return path
