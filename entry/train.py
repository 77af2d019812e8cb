#!/usr/bin/env python
"""Training entry point — the flags of the reference's train.py:23-41, checkpoints as train.py:47-53
(`epoch_model_XXX.pth`, epoch 0 included), Adam + MultiStepLR as train.py:332-340, per-snapshot validation PNGs and
`A_log.csv` as train.py:391-434.

Two loops (`--loop`):
  n2n         the Neighbor2Neighbor iteration of training_script.md:128-156 (the path BASELINE.json names; the fork's
              train.py carries its ingredients at :134-190 but no longer calls them) through the fused
              N2NTrainer: clean images only, noise added on the device;
  supervised  the fork's live loop, train.py:354-368: network(noisy), network(clean) with grad, util.Structure_loss,
              on <data_dir>/clean + <data_dir>/noise pairs.
The network family is picked from --log_name as train.py:298-314 does ('UNET' / 'RESNET' / 'UNetImproved'; under --loop n2n the
UNet runs the fused trainer, the other families the same iteration through autograd).  Multi-GPU:
`torchrun --nproc-per-node N entry/train.py --parallel ...` (one process per GPU, NCCL) replaces nn.DataParallel
(train.py:324-325).  Patches are cut on the device from images uploaded once (image_denoising_b200.data).

    python entry/train.py --data_dir data --log_name UNET_gauss25 --n_epoch 100 --batchsize 4
    python entry/train.py --synthetic 8 --n_epoch 2 --batchsize 64                  # no dataset needed
"""
import argparse
import datetime
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from entry import _data, _models  # noqa: E402
from image_denoising_b200 import (AugmentNoise, FusedAdam, N2NTrainer, Structure_loss, UNet, checkpoint, dp,  # noqa: E402
                                   forward_pair, generate_mask_pair, generate_subimage_pair, n2n_loss, ops)
from image_denoising_b200.data import DevicePatchSource  # noqa: E402
from image_denoising_b200.optim import multistep_lr  # noqa: E402

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
parser.add_argument("--loop", default="n2n", choices=["n2n", "supervised"])
parser.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
parser.add_argument("--patch", type=int, default=256, help="random-crop size of the training patches")
parser.add_argument("--patches_per_image", type=int, default=16)
parser.add_argument("--synthetic", type=int, default=0, help="train on this many synthetic image pairs instead of --data_dir")


def validate(network, valid, epoch, opt, validation_path, dev):
    """train.py:391-430: whole-image forward of every validation image, clip(p*255+0.5) -> uint8; PNGs of image 0."""
    clean_imgs, noisy_imgs, clean_paths, noise_paths = valid
    os.makedirs(validation_path, exist_ok=True)
    for i in range(len(clean_imgs)):
        clean_name = os.path.basename(clean_paths[i]).split('.')[0]
        noise_name = os.path.basename(noise_paths[i]).split('.')[0]
        noisy_im = np.asarray(noisy_imgs[i], dtype=np.float32) / 255.0
        t = torch.from_numpy(noisy_im[None] if noisy_im.ndim == 2 else np.transpose(noisy_im, (2, 0, 1))).unsqueeze(0).to(dev)
        with torch.no_grad():
            prediction = network(t)
        pred255 = np.squeeze(ops.quantize_u8(prediction, 0.5).permute(0, 2, 3, 1).cpu().numpy())
        if i == 0 and epoch == opt.n_snapshot:
            _data.save_rgb(np.asarray(clean_imgs[i]).astype(np.uint8), os.path.join(validation_path, "{}_{:03d}-{:03d}_clean.png".format(clean_name, i, epoch)))
            _data.save_rgb(np.asarray(noisy_imgs[i]).astype(np.uint8), os.path.join(validation_path, "{}_{:03d}-{:03d}_noisy.png".format(noise_name, i, epoch)))
        if i == 0:
            _data.save_rgb(pred255, os.path.join(validation_path, "{}_{:03d}-{:03d}_denoised.png".format(noise_name, i, epoch)))


def main():
    opt, _ = parser.parse_known_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    systime = datetime.datetime.now().strftime('%Y-%m-%d-%H-%M')

    # ---- data: images as the reference holds them (float32 0..255, train.py:208-228), resident on the device ----
    if opt.synthetic:
        size = max(opt.patch, 256)
        clean_u8, noisy_u8 = _data.synthetic_images(opt.synthetic, size, size, opt.n_channel)
        clean = [c.astype(np.float32) for c in clean_u8]; noise = [n.astype(np.float32) for n in noisy_u8]
        clean_paths = noise_paths = [f"synthetic_{i:03d}.png" for i in range(opt.synthetic)]
    else:
        clean_paths, noise_paths = _data.list_pairs(opt.data_dir)
        clean = [_data.load_image(f) for f in clean_paths]; noise = [_data.load_image(f) for f in noise_paths]
    valid = (clean, noise, clean_paths, noise_paths)                      # train.py:292: validation = the training pairs
    source = DevicePatchSource(clean, noise if opt.loop == "supervised" else None, device=dev)
    rng = np.random.RandomState(1234 + rank)
    per_rank = opt.batchsize if not opt.parallel else max(opt.batchsize // world, 1)

    torch.manual_seed(0)
    network = _models.network_from_log_name(opt.log_name, opt.n_channel, opt.n_feature).to(dev).set_precision(opt.precision)
    fused_n2n = opt.loop == "n2n" and type(network) is UNet
    if opt.loop == "n2n":
        noise_adder = AugmentNoise(style=opt.noisetype, rank=rank, world=world)      # global-batch noise, this rank's slice
    if fused_n2n:
        trainer = N2NTrainer(network, lr=opt.lr, precision=opt.precision)            # the whole iteration as one CUDA graph
    else:
        # RESNET / ImprovedUNet under --loop n2n: the same iteration (training_script.md:137-156) written out on the drop-in
        # functions with autograd; the supervised loop below shares the optimiser / criterion
        if world > 1:
            # parameters_to_vector returns a COPY: broadcast it, then write rank 0's values back into the parameters
            flat = torch.nn.utils.parameters_to_vector(network.parameters()).detach()
            dp.broadcast_params(flat, 0)
            off = 0
            for prm in network.parameters():
                prm.data.copy_(flat[off:off + prm.numel()].view_as(prm))
                off += prm.numel()
        optimizer = FusedAdam(network.parameters(), lr=opt.lr)
        criterion = Structure_loss()                                                 # train.py:322
    if rank == 0:
        checkpoint(network, 0, "model", opt.save_model_path, opt.log_name, systime)  # train.py:343
    steps_per_epoch = max(len(source) * opt.patches_per_image // (per_rank * world), 1)
    print(f"rank {rank}/{world}: {len(source)} images, {steps_per_epoch} steps/epoch, batch {per_rank}/GPU, loop {opt.loop}")
    for epoch in range(1, opt.n_epoch + 1):
        epoch_st = time.time()
        lr = multistep_lr(opt.lr, epoch, opt.n_epoch, opt.gamma)                     # train.py:333-340, :375
        if rank == 0:
            print("LearningRate of Epoch {} = {}".format(epoch, lr))
        l1_loss = []
        for it in range(steps_per_epoch):
            st = time.time()
            sel = source.draw(rng.randint(0, len(source), size=per_rank), opt.patch, rng)
            clean_b, noisy_b = source.crop(sel, opt.patch)                           # /255 on the device (train.py:358)
            if opt.loop == "n2n":
                Lambda = epoch / opt.n_epoch * opt.increase_ratio                    # training_script.md:148
                noisy_b = noise_adder.add_train_noise(clean_b)
                if fused_n2n:
                    loss3 = trainer.step(noisy_b, Lambda, lr=lr)
                else:
                    for group in optimizer.param_groups:
                        group['lr'] = lr
                    optimizer.zero_grad()
                    mask1, mask2 = generate_mask_pair(noisy_b)                       # training_script.md:137-144
                    noisy_sub1, noisy_sub2 = generate_subimage_pair(noisy_b, mask1, mask2)
                    with torch.no_grad():
                        noisy_denoised = network(noisy_b)
                    den_sub1, den_sub2 = generate_subimage_pair(noisy_denoised, mask1, mask2)
                    loss, loss3 = n2n_loss(network(noisy_sub1), noisy_sub2, den_sub1, den_sub2, Lambda)   # :146-153
                    loss.backward()
                    if world > 1:
                        dp.allreduce_mean_grads(network.parameters())
                    optimizer.step()
                if it % 50 == 0:
                    l = loss3.tolist()
                    l1_loss.append(l[1])
                    if rank == 0:
                        print('{:04d} {:05d} Loss1={:.6f}, Lambda={}, Loss2={:.6f}, Loss_Full={:.6f}, Time={:.4f}'.format(
                            epoch, it, l[1], Lambda, l[2], l[0], time.time() - st))
            else:
                for group in optimizer.param_groups:
                    group['lr'] = lr
                optimizer.zero_grad()
                noisy_output, clean_ = forward_pair(network, noisy_b, clean_b)       # train.py:361, one batched pass
                loss = criterion(noisy_output, clean_, clean_b)
                loss.backward()
                if world > 1:
                    dp.allreduce_mean_grads(network.parameters())
                optimizer.step()
                if it % 50 == 0:
                    terms = criterion.last_terms.tolist()                            # [loss, L1(noisy_output, clean), TV, cst]
                    l1_loss.append(terms[1])
                    if rank == 0:
                        print('{:04d} {:05d} Loss1={:.6f}, Loss_Full={:.6f}, Time={:.4f}'.format(epoch, it, terms[1], terms[0], time.time() - st))
        train_time = time.time() - epoch_st
        mean_loss = float(np.mean(l1_loss)) if l1_loss else float("nan")
        if rank == 0:
            print(f'Training Time/Epoch:{train_time} \n Mean loss:{mean_loss}')
        if rank == 0 and (epoch % opt.n_snapshot == 0 or epoch == opt.n_epoch):
            eval_st = time.time()
            checkpoint(network, epoch, "model", opt.save_model_path, opt.log_name, systime)
            validation_path = os.path.join(opt.save_model_path, opt.log_name, systime, "validation")
            try:
                validate(network, valid, epoch, opt, validation_path, dev)
            except ValueError as e:            # e.g. image sizes that are not multiples of 32 (the reference UNet fails there too)
                os.makedirs(validation_path, exist_ok=True)
                print(f"validation skipped: {e}")
            with open(os.path.join(validation_path, "A_log.csv"), "a") as f:       # train.py:431-434
                f.writelines("epoch{}, loss_{}, train_time_{}\n".format(epoch, mean_loss, train_time))
            print(f'Evaluation Time/Epoch:{time.time() - eval_st}')
    if world > 1:
        # the captured step graph holds NCCL work: drop it before the communicator goes away (destroy_process_group otherwise
        # waits on it forever), line the ranks up, and never let a stuck teardown outlive a finished training run
        trainer = None
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        import threading
        t = threading.Timer(60.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
        t.cancel()


if __name__ == "__main__":
    main()
