#!/usr/bin/env python
"""Whole-image evaluation — flags and semantics of the reference's evaluation.py:13-114 (one forward
per image, clip(p*255+0.5) -> uint8, PSNR / SSIM / L1(pred, noisy) -> metrics.txt).  With --tiled it
follows evaluation_704.py:57-130 instead (352x352 tiles at stride 288, triangular blend, truncation).
Images are sharded over ranks under torchrun (no collective)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from entry import _data  # noqa: E402
from image_denoising_b200 import UNet  # noqa: E402
from image_denoising_b200.evaluate import denoise_tiled, denoise_whole  # noqa: E402
from image_denoising_b200.utils_eval import psnr_ssim_batch  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--data_dir', type=str, default='./dataset/m1', help='dataset dir')
parser.add_argument('--checkpoint', type=str, default=None, help='path to checkpoint .pth file')
parser.add_argument('--save_dir', type=str, default='./eval_results', help='directory to save denoised images')
parser.add_argument('--n_feature', type=int, default=48)
parser.add_argument('--n_channel', type=int, default=1)
parser.add_argument('--log_name', type=str, default='UNET')
parser.add_argument('--gpu_devices', default='0', type=str)
parser.add_argument('--tiled', action='store_true', help='evaluation_704.py semantics')
parser.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
parser.add_argument('--synthetic', type=int, default=0, help='evaluate this many synthetic 704x704 pairs instead of --data_dir')
# evaluation_adapter.py:17-44: base + output adapter loaded from an epoch_adapter_XXX.pth checkpoint
parser.add_argument('--adapter_ckpt', type=str, default=None, help='epoch_adapter_XXX.pth (base.* + adapter.* keys)')
parser.add_argument('--adapter_hidden', type=int, default=16)


def main():
    opt = parser.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.makedirs(opt.save_dir, exist_ok=True)
    network = UNet(in_nc=opt.n_channel, out_nc=opt.n_channel, n_feature=opt.n_feature)
    if opt.checkpoint:
        state = torch.load(opt.checkpoint, map_location="cpu")
        network.load_state_dict(state)                      # strict, as evaluation.py:52-53
    network = network.to(f"cuda:{local}").set_precision(opt.precision).eval()
    if opt.adapter_ckpt:
        # evaluation_adapter.py:59-69, :100-112: strict=False load, "module." prefixes stripped
        from image_denoising_b200 import DenoiserWithAdapter
        state = torch.load(opt.adapter_ckpt, map_location="cpu")
        if any(k.startswith("module.") for k in state.keys()):
            state = {k.replace("module.", "", 1): v for k, v in state.items()}
        model = DenoiserWithAdapter(network, in_channels=opt.n_channel, hidden_channels=opt.adapter_hidden,
                                    freeze_base=True, use_no_grad_for_base=True)
        missing, unexpected = model.load_state_dict(state, strict=False)
        if missing:
            print(f"[Warning] Missing keys when loading adapter model: {missing}")
        if unexpected:
            print(f"[Warning] Unexpected keys when loading adapter model: {unexpected}")
        network = model.to(f"cuda:{local}").set_precision(opt.precision).eval()
    if opt.synthetic:
        clean, noisy = _data.synthetic_images(opt.synthetic, 704, 704, opt.n_channel)
        names = [f"synthetic_{i:03d}.png" for i in range(opt.synthetic)]
    else:
        cf, nf = _data.list_pairs(opt.data_dir)
        clean = [_data.load_image(f).astype(np.uint8) for f in cf]
        noisy = [_data.load_image(f) for f in nf]
        names = [os.path.basename(f) for f in nf]
    idx = list(range(rank, len(noisy), world))              # image sharding, no collective
    lines, psnrs, ssims = [], [], []
    by_shape = {}
    for i in idx:
        by_shape.setdefault(np.asarray(noisy[i]).shape, []).append(i)
    for shape, ids in by_shape.items():
        for b0 in range(0, len(ids), 8):
            chunk = ids[b0:b0 + 8]
            run = denoise_tiled if opt.tiled else denoise_whole
            preds, l1 = run(network, [noisy[i] for i in chunk], device=f"cuda:{local}")
            res = psnr_ssim_batch(preds, [clean[i] for i in chunk], device=f"cuda:{local}")
            for i, pr, l, (ps, ss) in zip(chunk, preds, l1, res):
                _data.save_image(pr, os.path.join(opt.save_dir, os.path.splitext(names[i])[0] + "_denoised.png"))
                psnrs.append(ps); ssims.append(ss)
                lines.append(f"{names[i]}: PSNR={ps:.4f}, SSIM={ss:.6f}, L1_pred_noisy={l:.6f}")
    with open(os.path.join(opt.save_dir, f"metrics_rank{rank}.txt" if world > 1 else "metrics.txt"), "w") as f:
        f.write("\n".join(lines) + f"\nAVG: PSNR={np.mean(psnrs):.4f}, SSIM={np.mean(ssims):.6f}\n")
    print(f"rank {rank}: {len(idx)} images, PSNR {np.mean(psnrs):.3f} dB, SSIM {np.mean(ssims):.5f}")


if __name__ == "__main__":
    main()
