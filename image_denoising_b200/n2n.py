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


def _normal_out(shape, std, generator, device):
    """``torch.normal(mean=0.0, std=std[B,1,1,1], generator=g, out=noise[B,C,H,W])`` as the reference's pinned
    PyTorch 1.3 (README.md:6-9) executes it: fill ``noise`` with N(0,1) at ITS OWN shape, then ``mul_(std)``
    (per-pixel noise, per-sample sigma).  torch >= 2 instead resizes ``out`` to std's shape [B,1,1,1] (with a
    deprecation warning), which would turn train.py:84-94 into one constant offset per sample; the intended
    per-pixel semantics are kept here."""
    noise = torch.empty(shape, dtype=torch.float32, device=device)
    noise.normal_(mean=0.0, std=1.0, generator=generator)
    return noise.mul_(std)


class AugmentNoise(object):
    """train.py:64-131 — synthetic noise for training (torch RNG; data generation only)."""

    def __init__(self, style, rank: int = 0, world: int = 1):
        """``rank`` / ``world`` (not in the reference): data-parallel ranks draw the noise of the GLOBAL batch from the
        same counter-seeded generator and keep their own slice, so that W processes see exactly the noise one
        process would have added to the concatenated batch (and never W copies of one realisation)."""
        self.rank, self.world = int(rank), int(world)
        if style.startswith('gauss'):
            self.params = [float(p) / 255.0 for p in style.replace('gauss', '', 1).split('_')]
            self.style = "gauss_fix" if len(self.params) == 1 else "gauss_range"
        elif style.startswith('poisson'):
            self.params = [float(p) for p in style.replace('poisson', '', 1).split('_')]
            self.style = "poisson_fix" if len(self.params) == 1 else "poisson_range"
        else:
            raise ValueError(f"unknown noise style {style!r}")

    def _mine(self, t):
        n = t.shape[0] // self.world
        return t[self.rank * n:(self.rank + 1) * n]

    def add_train_noise(self, x):
        dev = x.device
        shape = (x.shape[0] * self.world,) + tuple(x.shape[1:])        # the global batch
        if self.style == "gauss_fix":
            std = self.params[0] * torch.ones((shape[0], 1, 1, 1), device=dev)
            noise = _normal_out(shape, std, get_generator(dev), dev)
            return x + self._mine(noise)
        if self.style == "gauss_range":
            min_std, max_std = self.params
            std = torch.rand(size=(shape[0], 1, 1, 1), device=dev) * (max_std - min_std) + min_std
            noise = _normal_out(shape, std, get_generator(dev), dev)
            return x + self._mine(noise)
        if self.world > 1:
            raise NotImplementedError("Poisson training noise depends on the data; draw it per rank (world=1)")
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


def draw_rd_idx(img: torch.Tensor, batch: int = None) -> torch.Tensor:
    """train.py:155-162: one randint(0, 8) per 2x2 cell from a fresh counter-seeded generator.
    ``batch`` overrides the batch dimension: data-parallel ranks draw the selector of the GLOBAL batch
    (same counter seed on every rank) and keep their slice, so that W ranks use exactly the masks one
    process would use on the concatenated batch."""
    n, c, h, w = img.shape
    if batch is not None:
        n = int(batch)
    cells = n * h // 2 * w // 2
    rd_idx = torch.empty(size=(cells,), dtype=torch.int64, device=img.device)     # randint(out=) writes every element
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


def forward_pair(network, a, b):
    """``network(a), network(b)`` (train.py:361: ``network(noisy), network(clean)``) as ONE call on the concatenated batch.
    Every layer of UNet / RESNET / ImprovedUNet is per-sample (no batch statistics: GroupNorm normalises within a sample), so
    the two halves of the result are what the two calls return; one pass of twice the batch halves the launches and fills
    the machine better (measured on one B200, live supervised step: UNet 16 x 256^2 6.4 -> 4.8 ms, ImprovedUNet 4 x 128^2
    16.6 -> 7.9 ms)."""
    out = network(torch.cat([a, b], dim=0))
    n = a.shape[0]
    return out[:n], out[n:]


def space_to_depth(x, block_size):
    """train.py:134-138: F.unfold(x, block_size, stride=block_size).view(n, c*bs**2, h//bs, w//bs) as one gather
    kernel (generate_subimages itself never materialises it)."""
    return ops.space_to_depth(x, block_size)


def checkpoint(net, epoch, name, save_model_path, log_name, systime=None):
    """train.py:47-53: <save_model_path>/<log_name>/<systime>/epoch_<name>_<epoch:03d>.pth."""
    systime = systime or datetime.datetime.now().strftime('%Y-%m-%d-%H-%M')
    d = os.path.join(save_model_path, log_name, systime)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, 'epoch_{}_{:03d}.pth'.format(name, epoch))
    torch.save(net.state_dict(), path)
    print('Checkpoint saved to {}'.format(path))
    return path
