#!/usr/bin/env python
"""Adapter finetuning — flags of the reference's finetune.py:23-81; frozen base UNet under no_grad +
OutputAdapter (adapter.py:5-67), random-crop patch dataset over the first five image pairs
(finetune.py:94-150), loss = L1 + lambda_grad * gradient loss (finetune.py:153-162, :283-285), Adam over
the adapter parameters only (finetune.py:260-263), whole wrapper state_dict saved as
`epoch_adapter_XXX.pth` (finetune.py:84-91)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from entry import _data  # noqa: E402
from image_denoising_b200 import DenoiserWithAdapter, FusedAdam, UNet, l1_grad_loss  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--data_dir', type=str, default=None)
parser.add_argument('--pretrained_ckpt', type=str, default=None)
parser.add_argument('--arch', type=str, default='UNet', choices=['UNet', 'RESNET', 'UNetImproved'])
parser.add_argument('--save_model_path', type=str, default='./results_ft')
parser.add_argument('--log_name', type=str, default='UNet_adapter_ft')
parser.add_argument('--gpu_devices', default='0', type=str)
parser.add_argument('--parallel', action='store_true')
parser.add_argument('--n_feature', type=int, default=48)
parser.add_argument('--n_channel', type=int, default=1)
parser.add_argument('--lr', type=float, default=1e-4)
parser.add_argument('--n_epoch', type=int, default=20)
parser.add_argument('--batchsize', type=int, default=4)
parser.add_argument('--num_workers', type=int, default=4)
parser.add_argument('--adapter_hidden', type=int, default=16)
parser.add_argument('--lambda_grad', type=float, default=0.1)
parser.add_argument('--save_every', type=int, default=1)
parser.add_argument('--patch_size', type=int, default=128)
parser.add_argument('--patches_per_image', type=int, default=16)
parser.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
parser.add_argument('--synthetic', type=int, default=0)


def main():
    args, _ = parser.parse_known_args()
    if args.arch != 'UNet':
        raise SystemExit("only --arch UNet is on the B200 path (SURVEY.md §8f)")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if args.synthetic:
        clean_u8, noisy_u8 = _data.synthetic_images(args.synthetic, 256, 256, args.n_channel)
        clean = [c.astype(np.float32) for c in clean_u8]; noise = [n.astype(np.float32) for n in noisy_u8]
    else:
        cf, nf = _data.list_pairs(args.data_dir, limit=5)           # finetune.py:109-110
        clean = [_data.load_image(f) for f in cf]; noise = [_data.load_image(f) for f in nf]
    chw = lambda a: a[None] if a.ndim == 2 else np.transpose(a, (2, 0, 1))
    clean = [chw(a) for a in clean]; noise = [chw(a) for a in noise]

    base = UNet(in_nc=args.n_channel, out_nc=args.n_channel, n_feature=args.n_feature)
    if args.pretrained_ckpt:
        state = torch.load(args.pretrained_ckpt, map_location="cpu")
        state = {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}     # finetune.py:207-218
        base.load_state_dict(state, strict=False)
    model = DenoiserWithAdapter(base, in_channels=args.n_channel, hidden_channels=args.adapter_hidden,
                                freeze_base=True, use_no_grad_for_base=True).to(dev)
    model.set_precision(args.precision)
    opt = FusedAdam(filter(lambda p: p.requires_grad, model.parameters()), lr=args.lr)
    out_dir = os.path.join(args.save_model_path, args.log_name)
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(0)
    ps = args.patch_size
    samples = len(clean) * args.patches_per_image
    for epoch in range(1, args.n_epoch + 1):
        order = rng.permutation(samples)
        tot = 0.0
        for b0 in range(0, samples - args.batchsize + 1, args.batchsize):
            cb = np.empty((args.batchsize, args.n_channel, ps, ps), np.float32); nb = np.empty_like(cb)
            for j, s in enumerate(order[b0:b0 + args.batchsize]):
                i = s // args.patches_per_image
                top = rng.integers(0, clean[i].shape[1] - ps + 1); left = rng.integers(0, clean[i].shape[2] - ps + 1)
                cb[j] = clean[i][:, top:top + ps, left:left + ps]; nb[j] = noise[i][:, top:top + ps, left:left + ps]
            c = torch.from_numpy(cb).to(dev) / 255.0; n = torch.from_numpy(nb).to(dev) / 255.0
            opt.zero_grad(set_to_none=True)
            loss, _loss3 = l1_grad_loss(model(n), c, args.lambda_grad)
            loss.backward()
            opt.step()
            tot += float(loss)
        print(f"[Epoch {epoch:03d}] loss {tot / max(samples // args.batchsize, 1):.6f}")
        if epoch % args.save_every == 0 or epoch == args.n_epoch:
            path = os.path.join(out_dir, 'epoch_adapter_{:03d}.pth'.format(epoch))
            torch.save(model.state_dict(), path)
            print('Checkpoint saved to {}'.format(path))


if __name__ == "__main__":
    main()
