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
from entry import _data, _models  # noqa: E402
from image_denoising_b200 import DenoiserWithAdapter, FusedAdam, iqsl_loss, l1_grad_loss, ops  # noqa: E402
from image_denoising_b200.data import DevicePatchSource  # noqa: E402

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


def main(args=None, iqsl=False):
    """``iqsl`` = the finetune_iqsl.py variant of the same loop (entry/finetune_iqsl.py): adds lambda_iqsl * iqsl_loss
    (finetune_iqsl.py:469-483) with thresholds estimated from the clean images (:258-288) and saves the ADAPTER's
    state_dict only as `epoch_adapter_only_XXX.pth` (:114-132)."""
    if args is None:
        args, _ = parser.parse_known_args()
    import datetime
    systime = datetime.datetime.now().strftime('%Y-%m-%d-%H-%M')
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if args.synthetic:
        clean_u8, noisy_u8 = _data.synthetic_images(args.synthetic, 256, 256, args.n_channel)
        clean = [c.astype(np.float32) for c in clean_u8]; noise = [n.astype(np.float32) for n in noisy_u8]
    else:
        cf, nf = _data.list_pairs(args.data_dir, limit=5)           # finetune.py:109-110
        clean = [_data.load_image(f) for f in cf]; noise = [_data.load_image(f) for f in nf]
    # the (up to five) image pairs live on the device; patches are cut there (image_denoising_b200.data)
    source = DevicePatchSource(clean, noise, device=dev)
    valid_clean, valid_noise = clean, noise                       # finetune.py:231: validation = the same pairs, whole images
    t1 = t2 = None
    if iqsl and args.lambda_iqsl > 0.0:
        # finetune_iqsl.py:258-288: quantiles of the pooled clean pixels in [0,1] (all clean files up to iqsl_max_images)
        if args.synthetic:
            pool = clean
        else:
            pool = [_data.load_image(f) for f in _data.list_pairs(args.data_dir, limit=args.iqsl_max_images)[0]]
        assert 0.0 < args.iqsl_q1 < args.iqsl_q2 < 1.0, 'iqsl_q1, iqsl_q2 must satisfy 0 < q1 < q2 < 1.'
        px = np.concatenate([np.asarray(c, np.float32).reshape(-1) / 255.0 for c in pool])
        t1, t2 = float(np.quantile(px, args.iqsl_q1)), float(np.quantile(px, args.iqsl_q2))
        print(f'[IQSL] Estimated thresholds from clean/: t1={t1:.4f}, t2={t2:.4f}')
    elif iqsl:
        print('[IQSL] lambda_iqsl=0 → IQSL disabled.')

    base = _models.build_base_model(args.arch, args.n_channel, args.n_feature)                # finetune.py:189-204
    if args.pretrained_ckpt:
        state = _models.strip_module_prefix(torch.load(args.pretrained_ckpt, map_location="cpu"))   # finetune.py:207-218
        missing, unexpected = base.load_state_dict(state, strict=False)
        if missing:
            print(f'[Warning] Missing keys when loading base model: {missing}')
        if unexpected:
            print(f'[Warning] Unexpected keys when loading base model: {unexpected}')
    model = DenoiserWithAdapter(base, in_channels=args.n_channel, hidden_channels=args.adapter_hidden,
                                freeze_base=True, use_no_grad_for_base=True).to(dev)
    model.set_precision(args.precision)
    opt = FusedAdam(filter(lambda p: p.requires_grad, model.parameters()), lr=args.lr)
    out_dir = os.path.join(args.save_model_path, args.log_name)
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.RandomState(0)
    ps = args.patch_size
    samples = len(source) * args.patches_per_image               # finetune.py:123-124
    for epoch in range(1, args.n_epoch + 1):
        order = rng.permutation(samples)                           # DataLoader(shuffle=True, drop_last=False), finetune.py:224-231
        losses = []
        for it, b0 in enumerate(range(0, samples, args.batchsize)):
            sel = source.draw([int(s) // args.patches_per_image for s in order[b0:b0 + args.batchsize]], ps, rng)
            c, n = source.crop(sel, ps)                            # same (top, left) for clean and noise, /255 (finetune.py:136-147)
            opt.zero_grad(set_to_none=True)
            pred = model(n)
            loss, loss3 = l1_grad_loss(pred, c, args.lambda_grad)
            loss_iq = None
            if t1 is not None:
                loss_iq = iqsl_loss(pred, c, t1=t1, t2=t2, tau=args.iqsl_tau, margin=args.iqsl_margin,
                                    ce_factor=args.iqsl_ce_factor)
                loss = loss + args.lambda_iqsl * loss_iq                # finetune_iqsl.py:483
            loss.backward()
            opt.step()
            if it % 10 == 0:
                l = loss3.tolist()
                losses.append(l[1])
                if loss_iq is not None:
                    print(f'[Epoch {epoch:03d} | Iter {it:04d}] L1={l[1]:.6f} Grad={l[2]:.6f} IQSL={loss_iq.item():.6f} '
                          f'Total={loss.item():.6f}')
                else:
                    print(f'[Epoch {epoch:03d} | Iter {it:04d}] L1={l[1]:.6f} Grad={l[2]:.6f} Total={l[0]:.6f}')
        print(f'End of epoch {epoch}, mean L1 loss={float(np.mean(losses)):.6f}')
        if epoch % args.save_every == 0 or epoch == args.n_epoch:
            if iqsl:                                                   # finetune_iqsl.py:114-132: the adapter alone
                path = os.path.join(out_dir, 'epoch_adapter_only_{:03d}.pth'.format(epoch))
                torch.save(model.adapter.state_dict(), path)
                print('Adapter checkpoint saved to {}'.format(path))
            else:
                path = os.path.join(out_dir, 'epoch_adapter_{:03d}.pth'.format(epoch))
                torch.save(model.state_dict(), path)
                print('Checkpoint saved to {}'.format(path))
            # finetune.py:305-343: whole-image validation PSNR (clip(p*255+0.5), 99.0 when identical), PNGs of image 0
            save_dir = os.path.join(args.save_model_path, args.log_name, f'val_{systime}_ep{epoch:03d}')
            os.makedirs(save_dir, exist_ok=True)
            model.eval()
            psnrs = []
            with torch.no_grad():
                for i, (clean_np, noisy_np) in enumerate(zip(valid_clean, valid_noise)):
                    noisy_im = np.asarray(noisy_np, np.float32) / 255.0
                    t = torch.from_numpy(noisy_im[None] if noisy_im.ndim == 2 else np.transpose(noisy_im, (2, 0, 1))).unsqueeze(0).to(dev)
                    try:
                        pred = model(t)
                    except ValueError as e:
                        print(f"validation skipped: {e}")
                        break
                    pred255 = np.squeeze(ops.quantize_u8(pred, 0.5).squeeze(0).permute(1, 2, 0).cpu().numpy())
                    diff = pred255.astype(np.float32) - np.asarray(clean_np, np.float32)
                    mse = float(np.mean(np.square(diff)))
                    psnrs.append(99.0 if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse))
                    if i == 0:
                        _data.save_image(np.asarray(clean_np).astype(np.uint8), os.path.join(save_dir, f'clean_{i:03d}.png'))
                        _data.save_image(np.asarray(noisy_np).astype(np.uint8), os.path.join(save_dir, f'noisy_{i:03d}.png'))
                        _data.save_image(pred255, os.path.join(save_dir, f'denoised_ep{epoch:03d}.png'))
            if psnrs:
                print(f'Val ep{epoch}: PSNR {float(np.mean(psnrs)):.2f} dB over {len(psnrs)} images')
            model.train()
    print('Finetuning complete.')


if __name__ == "__main__":
    main()
