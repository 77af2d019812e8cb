"""Device-side data path: the reference crops patches on the host inside DataLoader workers
(train.py:208-228 — whole images; finetune.py:94-150 — ``DenoisePatchDataset``: random ps x ps crops at identical
coordinates of a clean / noisy pair, ``/255``) and ships every batch over PCIe.  ``DevicePatchSource`` uploads the
images ONCE (float32 H x W x C, 0..255, as the reference holds them) and cuts each batch with one n2n_crop_patches
launch per tensor; only the 12 B/patch of crop coordinates cross PCIe per step.  The coordinates come from numpy's
RNG exactly as the reference draws them (``np.random.randint(0, h - ps + 1)``, finetune.py:139-140)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._ext import check, lib, ptr, require_cuda, stream_ptr


def _hwc(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a, dtype=np.float32)
    return np.ascontiguousarray(a[:, :, None] if a.ndim == 2 else a)


class DevicePatchSource:
    def __init__(self, clean: Sequence[np.ndarray], noise: Optional[Sequence[np.ndarray]] = None, device="cuda"):
        self.device = torch.device(device)
        self.clean = [torch.from_numpy(_hwc(c)).to(self.device) for c in clean]
        self.noise = [torch.from_numpy(_hwc(n)).to(self.device) for n in noise] if noise is not None else None
        require_cuda(self.clean[0], "DevicePatchSource")
        self.channels = int(self.clean[0].shape[2])
        if noise is not None:
            assert len(self.noise) == len(self.clean) and all(a.shape == b.shape for a, b in zip(self.clean, self.noise))
        self.dims = torch.tensor([[t.shape[0], t.shape[1]] for t in self.clean], dtype=torch.int32, device=self.device)
        self._tab_c = torch.tensor([t.data_ptr() for t in self.clean], dtype=torch.int64, device=self.device)
        self._tab_n = (torch.tensor([t.data_ptr() for t in self.noise], dtype=torch.int64, device=self.device)
                       if self.noise is not None else None)
        self.shapes = [(int(t.shape[0]), int(t.shape[1])) for t in self.clean]

    def __len__(self):
        return len(self.clean)

    def draw(self, image_indices: Sequence[int], patch: int, rng=np.random) -> np.ndarray:
        """finetune.py:136-140: one (top, left) per sample, uniform over the valid range of its image."""
        sel = np.empty((len(image_indices), 3), np.int32)
        for j, i in enumerate(image_indices):
            h, w = self.shapes[i]
            if h < patch or w < patch:
                raise ValueError(f"Image size ({h},{w}) smaller than patch_size {patch}.")
            sel[j] = (i, rng.randint(0, h - patch + 1), rng.randint(0, w - patch + 1))
        return sel

    def crop(self, sel: np.ndarray, patch: int, scale: float = 1.0 / 255.0, out: Optional[Tuple[torch.Tensor, ...]] = None):
        """-> (clean [B,C,ps,ps], noisy [B,C,ps,ps] or None), fp32, scaled (the reference divides by 255 after ToTensor)."""
        sel = np.ascontiguousarray(sel, dtype=np.int32)
        b = sel.shape[0]
        for i, t, l in sel:
            h, w = self.shapes[int(i)]
            if t < 0 or l < 0 or t + patch > h or l + patch > w:
                raise ValueError("crop window outside the image")
        sel_d = torch.from_numpy(sel).to(self.device, non_blocking=True)
        outs: List[Optional[torch.Tensor]] = []
        for k, tab in enumerate((self._tab_c, self._tab_n)):
            if tab is None:
                outs.append(None)
                continue
            o = out[k] if out is not None else torch.empty((b, self.channels, patch, patch), dtype=torch.float32, device=self.device)
            check(lib().n2n_crop_patches(ptr(tab), ptr(self.dims), ptr(sel_d), b, self.channels, patch, float(scale), ptr(o),
                                         stream_ptr()))
            outs.append(o)
        return outs[0], outs[1]
