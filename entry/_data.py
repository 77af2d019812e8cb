"""Shared data helpers for the entry points: the reference's directory conventions
(<data_dir>/clean/*, <data_dir>/noise/*, matched by sorted file name — train.py:208-228,
finetune.py:94-150, utils_eval.py:7-17) plus a synthetic source for machines without the datasets."""
from __future__ import annotations

import glob
import os

import numpy as np


def list_pairs(data_dir: str, limit=None):
    clean = sorted(glob.glob(os.path.join(data_dir, "clean", "*")))
    noise = sorted(glob.glob(os.path.join(data_dir, "noise", "*")))
    if limit:
        clean, noise = clean[:limit], noise[:limit]
    if len(clean) != len(noise) or not clean:
        raise FileNotFoundError(f"{data_dir}: clean/ and noise/ must hold the same (non-zero) number of images")
    return clean, noise


def load_image(path: str) -> np.ndarray:
    from PIL import Image
    return np.array(Image.open(path), dtype=np.float32)       # 0..255 float32, as the reference loads them


def save_image(arr_u8: np.ndarray, path: str) -> None:
    from PIL import Image
    Image.fromarray(np.squeeze(arr_u8)).save(path)


def save_rgb(arr_u8: np.ndarray, path: str) -> None:
    """train.py:413-430: Image.fromarray(x).convert('RGB').save(path)."""
    from PIL import Image
    Image.fromarray(np.squeeze(arr_u8)).convert('RGB').save(path)


def synthetic_images(count: int, h: int, w: int, channels: int, sigma: float = 25.0, seed: int = 2025):
    """SEM-like smooth random fields + Gaussian noise, uint8."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    clean, noisy = [], []
    for _ in range(count):
        planes = []
        for _c in range(channels):
            f = gaussian_filter(rng.random((h, w)), 8.0)
            planes.append((f - f.min()) / max(f.max() - f.min(), 1e-12) * 255.0)
        c = np.stack(planes, -1).squeeze(-1) if channels == 1 else np.stack(planes, -1)
        n = np.clip(c + rng.normal(0.0, sigma, c.shape), 0, 255)
        clean.append(c.astype(np.uint8)); noisy.append(n.astype(np.uint8))
    return clean, noisy
