#!/usr/bin/env python
"""Adapter inference — the flags and semantics of the reference's evaluation_adapter.py:17-166: base network
(`--arch`) + OutputAdapter loaded from an `epoch_adapter_XXX.pth` checkpoint of the whole wrapper (strict=False,
`module.` prefixes stripped, :59-69), one whole-image forward per noisy image, `clip(p*255+0.5)` -> uint8,
`<name>_denoised.png`, optional PSNR (99.0 when identical, :72-80) when `<data_dir>/clean` exists."""
import argparse
import glob
import os
import sys

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from entry import _models  # noqa: E402
from image_denoising_b200 import DenoiserWithAdapter, ops  # noqa: E402


def parse_args():
    parser = argparse.ArgumentParser()
    parser.add_argument('--data_dir', type=str, required=True, help='Root dir with noise/ (and optionally clean/) for inference.')
    parser.add_argument('--ckpt', type=str, required=True, help='Checkpoint of DenoiserWithAdapter (epoch_adapter_xxx.pth).')
    parser.add_argument('--arch', type=str, default='UNetImproved', choices=['UNet', 'RESNET', 'UNetImproved'],
                        help='Backbone architecture used in base model.')
    parser.add_argument('--save_dir', type=str, default='./results_infer_adapter', help='Directory to save denoised images.')
    parser.add_argument('--gpu_devices', default='0', type=str)
    parser.add_argument('--parallel', action='store_true')
    parser.add_argument('--n_feature', type=int, default=48)
    parser.add_argument('--n_channel', type=int, default=1)
    parser.add_argument('--adapter_hidden', type=int, default=16)
    parser.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])      # addition
    args, _ = parser.parse_known_args()
    return args


def load_adapter_weights(model, ckpt_path: str):
    """evaluation_adapter.py:59-69."""
    state = _models.strip_module_prefix(torch.load(ckpt_path, map_location='cpu'))
    missing, unexpected = model.load_state_dict(state, strict=False)
    if missing:
        print(f'[Warning] Missing keys when loading adapter model: {missing}')
    if unexpected:
        print(f'[Warning] Unexpected keys when loading adapter model: {unexpected}')
    print(f'Loaded adapter model weights from {ckpt_path}')


def calculate_psnr(target: np.ndarray, ref: np.ndarray) -> float:
    """evaluation_adapter.py:72-80 (mse == 0 -> 99.0)."""
    diff = target.astype(np.float32) - ref.astype(np.float32)
    mse = np.mean(np.square(diff))
    if mse == 0:
        return 99.0
    return float(10.0 * np.log10(255.0 * 255.0 / mse))


def main():
    opt = parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    noise_dir = os.path.join(opt.data_dir, 'noise')
    clean_dir = os.path.join(opt.data_dir, 'clean')
    os.makedirs(opt.save_dir, exist_ok=True)
    noise_paths = sorted(glob.glob(os.path.join(noise_dir, '*')))
    if len(noise_paths) == 0:
        raise RuntimeError(f'No files found in {noise_dir}')
    has_clean = os.path.isdir(clean_dir) and len(glob.glob(os.path.join(clean_dir, '*'))) > 0
    clean_paths = []
    if has_clean:
        clean_paths = sorted(glob.glob(os.path.join(clean_dir, '*')))
        if len(clean_paths) != len(noise_paths):
            print('[Warning] clean/ and noise/ have different counts; PSNR may be misaligned.')
    print(f'Found {len(noise_paths)} noisy images for inference.')
    base_model = _models.build_base_model(opt.arch, opt.n_channel, opt.n_feature)
    model = DenoiserWithAdapter(base_model=base_model, in_channels=opt.n_channel, hidden_channels=opt.adapter_hidden,
                                freeze_base=True, use_no_grad_for_base=True)
    # --parallel (nn.DataParallel in the reference) = one process per GPU under torchrun: images are sharded
    load_adapter_weights(model, opt.ckpt)
    model = model.to(device).set_precision(opt.precision).eval()
    with torch.no_grad():
        for idx in range(rank, len(noise_paths), world):
            n_path = noise_paths[idx]
            name = os.path.basename(n_path)
            base_name = os.path.splitext(name)[0]
            noisy_norm = np.array(Image.open(n_path), dtype=np.float32) / 255.0
            t = torch.from_numpy(noisy_norm[None] if noisy_norm.ndim == 2 else np.transpose(noisy_norm, (2, 0, 1)))
            pred = model(t.unsqueeze(0).to(device))
            # evaluation_adapter.py:137-143: no clamp before the quantiser there; np.clip(p*255+0.5, 0, 255) equals
            # clamp(p, 0, 1) followed by the same expression for every p, so the fused quantiser is exact
            pred255 = ops.quantize_u8(pred, 0.5).squeeze(0).permute(1, 2, 0).cpu().numpy()
            if pred255.shape[2] == 1:
                out_img = Image.fromarray(pred255.squeeze(-1)).convert('L')
            else:
                out_img = Image.fromarray(pred255).convert('RGB')
            save_path = os.path.join(opt.save_dir, f'{base_name}_denoised.png')
            out_img.save(save_path)
            if has_clean and idx < len(clean_paths):
                clean_img = np.array(Image.open(clean_paths[idx]), dtype=np.float32)
                psnr = calculate_psnr(np.squeeze(pred255), clean_img)
                print(f'[{idx+1:03d}/{len(noise_paths):03d}] {name} → PSNR={psnr:.2f} dB, saved to {save_path}')
            else:
                print(f'[{idx+1:03d}/{len(noise_paths):03d}] {name} → saved to {save_path}')
    print('Inference with adapter model finished.')


if __name__ == '__main__':
    main()
