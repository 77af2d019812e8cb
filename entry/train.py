#!/usr/bin/env python
"""Neighbor2Neighbor training entry point — the flags of the reference's train.py:23-41, the loop of
training_script.md:128-156 (the fork's train.py carries the ingredients but runs a supervised loop,
SURVEY.md §0.2), checkpoints as train.py:47-53 (`epoch_model_XXX.pth`, epoch 0 included), Adam +
MultiStepLR as train.py:332-340.  Multi-GPU: `torchrun --nproc-per-node N entry/train.py --parallel ...`
(one process per GPU, NCCL) replaces the reference's nn.DataParallel (train.py:324-325).

    python entry/train.py --data_dir data --log_name UNET_gauss25 --n_epoch 100 --batchsize 4
    python entry/train.py --synthetic 256 --n_epoch 2 --batchsize 64       # no dataset needed
"""
import argparse
import datetime
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from entry import _data  # noqa: E402
from image_denoising_b200 import AugmentNoise, N2NTrainer, UNet, checkpoint, dp  # noqa: E402
from image_denoising_b200.optim import multistep_lr  # noqa: E402
from image_denoising_b200.prefetch import DevicePrefetcher  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--noisetype", type=str, default="gauss25")
parser.add_argument('--data_dir', type=str, default='data')
parser.add_argument('--save_model_path', type=str, default='./results')
parser.add_argument('--log_name', type=str, default='UNET_gauss25_b4e100r02')
parser.add_argument('--gpu_devices', default='0', type=str)
parser.add_argument('--parallel', action='store_true')
parser.add_argument('--n_feature', type=int, default=48)
parser.add_argument('--n_channel', type=int, default=1)
parser.add_argument('--lr', type=float, default=3e-4)
parser.add_argument('--gamma', type=float, default=0.5)
parser.add_argument('--n_epoch', type=int, default=100)
parser.add_argument('--n_snapshot', type=int, default=1)
parser.add_argument('--batchsize', type=int, default=4)
parser.add_argument("--Lambda1", type=float, default=1.0)
parser.add_argument("--Lambda2", type=float, default=1.0)
parser.add_argument("--increase_ratio", type=float, default=2.0)
# additions (not in the reference)
parser.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
parser.add_argument("--patch", type=int, default=256, help="random-crop size of the training patches")
parser.add_argument("--synthetic", type=int, default=0, help="train on this many synthetic clean images instead of --data_dir")


def main():
    opt, _ = parser.parse_known_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if 'UNET' not in opt.log_name and 'unet' not in opt.log_name.lower():
        raise SystemExit("only the UNet family (log_name containing 'UNET', train.py:298-314) is on the B200 path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    systime = datetime.datetime.now().strftime('%Y-%m-%d-%H-%M')

    # ---- data: clean images in 0..255 float32 (train.py:208-228); noise is added on the device ----
    if opt.synthetic:
        clean_u8, _ = _data.synthetic_images(opt.synthetic, max(opt.patch, 256), max(opt.patch, 256), opt.n_channel)
        images = [c.astype(np.float32) for c in clean_u8]
    else:
        clean_files, _ = _data.list_pairs(opt.data_dir)
        images = [_data.load_image(f) for f in clean_files]
    images = [im[None] if im.ndim == 2 else np.transpose(im, (2, 0, 1)) for im in images]
    rng = np.random.default_rng(1234 + rank)
    per_rank = opt.batchsize if not opt.parallel else max(opt.batchsize // world, 1)

    def next_batch():
        out = np.empty((per_rank, opt.n_channel, opt.patch, opt.patch), np.float32)
        for i in range(per_rank):
            im = images[rng.integers(len(images))]
            top = rng.integers(0, im.shape[1] - opt.patch + 1); left = rng.integers(0, im.shape[2] - opt.patch + 1)
            out[i] = im[:, top:top + opt.patch, left:left + opt.patch]
        return torch.from_numpy(out).pin_memory()

    torch.manual_seed(0)
    network = UNet(in_nc=opt.n_channel, out_nc=opt.n_channel, n_feature=opt.n_feature).to(dev).set_precision(opt.precision)
    noise_adder = AugmentNoise(style=opt.noisetype, rank=rank, world=world)      # global-batch noise, this rank's slice
    trainer = N2NTrainer(network, lr=opt.lr, precision=opt.precision)
    if rank == 0:
        checkpoint(network, 0, "model", opt.save_model_path, opt.log_name, systime)      # train.py:343
    staged = DevicePrefetcher(torch.empty((per_rank, opt.n_channel, opt.patch, opt.patch), dtype=torch.float32, device=dev))
    staged.put(next_batch())
    steps_per_epoch = max(len(images) * 16 // (per_rank * world), 1)
    print(f"rank {rank}/{world}: {len(images)} images, {steps_per_epoch} steps/epoch, batch {per_rank}/GPU")
    for epoch in range(1, opt.n_epoch + 1):
        lr = multistep_lr(opt.lr, epoch, opt.n_epoch, opt.gamma)      # train.py:333-340, :375
        Lambda = epoch / opt.n_epoch * opt.increase_ratio                                   # training_script.md:148
        st = time.time()
        for it in range(steps_per_epoch):
            # batch i+1 is cropped on the host and copied H2D (copy stream) while step i runs
            clean = staged.get() / 255.0
            staged.release()
            staged.put(next_batch())
            noisy = noise_adder.add_train_noise(clean)
            loss3 = trainer.step(noisy, Lambda, lr=lr)
            if it % 50 == 0 and rank == 0:
                l = loss3.tolist()
                print('{:04d} {:05d} Loss1={:.6f}, Lambda={}, Loss2={:.6f}, Loss_Full={:.6f}, Time={:.4f}'.format(
                    epoch, it, l[1], Lambda, l[2], l[0], time.time() - st))
        if rank == 0 and (epoch % opt.n_snapshot == 0 or epoch == opt.n_epoch):
            checkpoint(network, epoch, "model", opt.save_model_path, opt.log_name, systime)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
