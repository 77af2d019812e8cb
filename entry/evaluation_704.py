#!/usr/bin/env python
"""Tiled 704x704 evaluation — the flags, file names and semantics of the reference's evaluation_704.py:12-146:
352x352 tiles at stride 288 (edge tiles reflect-padded), triangular blend whose border weight is exactly 0,
`clip(p*255)` -> uint8 (truncation, no +0.5), PSNR / SSIM / L1(pred tile, noisy tile), `metrics.txt` with the three
averages.  The nine tiles of every image (and several images) run as ONE batched forward on the B200 engine; images
are sharded over ranks under torchrun (no collective)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from entry import _data, _models  # noqa: E402
from image_denoising_b200.evaluate import denoise_tiled  # noqa: E402
from image_denoising_b200.utils_eval import psnr_ssim_batch, validation_denoise  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--data_dir', type=str, default='./dataset/m1', help='dataset dir')
parser.add_argument('--checkpoint', type=str, required=True, help='path to checkpoint .pth file')
parser.add_argument('--save_dir', type=str, default='./eval_results', help='directory to save denoised images')
parser.add_argument('--n_feature', type=int, default=48)
parser.add_argument('--n_channel', type=int, default=1)
parser.add_argument('--log_name', type=str, default='UNetImproved')
parser.add_argument('--gpu_devices', default='0', type=str)
# additions (not in the reference)
parser.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
parser.add_argument('--synthetic', type=int, default=0, help='evaluate this many synthetic 704x704 pairs instead of --data_dir')


def evaluate():
    opt = parser.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    os.makedirs(opt.save_dir, exist_ok=True)
    if opt.synthetic:
        clean_imgs, noisy_imgs = _data.synthetic_images(opt.synthetic, 704, 704, opt.n_channel)
        clean_paths = noisy_paths = [f"synthetic_{i:03d}.png" for i in range(opt.synthetic)]
    else:
        clean_imgs, noisy_imgs, clean_paths, noisy_paths = validation_denoise(opt.data_dir)
    network = _models.network_from_log_name(opt.log_name, opt.n_channel, opt.n_feature)
    if opt.checkpoint != "random":
        network.load_state_dict(torch.load(opt.checkpoint, map_location="cpu"))      # strict, evaluation_704.py:46-47
        print(f"Loaded checkpoint from {opt.checkpoint}")
    network = network.to(dev).set_precision(opt.precision).eval()
    psnr_list, ssim_list, l1_list = [], [], []
    idx = list(range(rank, len(noisy_imgs), world))
    by_shape = {}
    for i in idx:
        by_shape.setdefault(np.asarray(noisy_imgs[i]).shape, []).append(i)
    for shape, ids in by_shape.items():
        for b0 in range(0, len(ids), 8):
            chunk = ids[b0:b0 + 8]
            noisy_u8 = [np.asarray(noisy_imgs[i]).astype(np.uint8) for i in chunk]      # evaluation_704.py:79-80
            clean_u8 = [np.asarray(clean_imgs[i]).astype(np.uint8) for i in chunk]
            preds, l1 = denoise_tiled(network, noisy_u8, device=dev)
            res = psnr_ssim_batch(preds, clean_u8, device=dev)
            for i, nz, cl, pr, l, (ps, ss) in zip(chunk, noisy_u8, clean_u8, preds, l1, res):
                clean_name = os.path.basename(clean_paths[i]).split('.')[0]
                noisy_name = os.path.basename(noisy_paths[i]).split('.')[0]
                _data.save_image(nz, os.path.join(opt.save_dir, f"{noisy_name}_{i:03d}_noisy.png"))
                _data.save_image(cl, os.path.join(opt.save_dir, f"{clean_name}_{i:03d}_clean.png"))
                _data.save_image(pr, os.path.join(opt.save_dir, f"{noisy_name}_{i:03d}_denoised.png"))
                psnr_list.append(ps); ssim_list.append(ss); l1_list.append(l)
                print(f"[{i+1}/{len(clean_imgs)}] {noisy_name} -> PSNR: {ps:.2f}, SSIM: {ss:.4f}, L1: {l:.6f}")
    avg_psnr, avg_ssim, avg_l1 = np.mean(psnr_list), np.mean(ssim_list), np.mean(l1_list)
    log_path = os.path.join(opt.save_dir, f"metrics_rank{rank}.txt" if world > 1 else "metrics.txt")
    with open(log_path, "w") as f:
        f.write(f"Average PSNR: {avg_psnr:.2f}\n")
        f.write(f"Average SSIM: {avg_ssim:.4f}\n")
        f.write(f"Average L1 Loss: {avg_l1:.6f}\n")
    print(f"Saved metrics to {log_path}")
    print(f"Average PSNR: {avg_psnr:.2f}, Average SSIM: {avg_ssim:.4f}, Average L1 Loss: {avg_l1:.6f}")


if __name__ == "__main__":
    evaluate()
