"""The reference-flag entry points (entry/train.py, entry/finetune.py, entry/evaluation.py) run end to end
on synthetic data and write the reference's checkpoint / metrics file names."""
import glob
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(args, cwd):
    r = subprocess.run([sys.executable] + args, cwd=cwd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def test_train_finetune_eval_roundtrip(tmp_path):
    out = str(tmp_path)
    _run([os.path.join(ROOT, "entry", "train.py"), "--synthetic", "4", "--patch", "64", "--batchsize", "2", "--n_epoch", "2",
          "--save_model_path", out, "--log_name", "UNET_test"], ROOT)
    ckpts = sorted(glob.glob(os.path.join(out, "UNET_test", "*", "epoch_model_*.pth")))
    assert [os.path.basename(c) for c in ckpts] == ["epoch_model_000.pth", "epoch_model_001.pth", "epoch_model_002.pth"]
    sd = torch.load(ckpts[-1], map_location="cpu")
    assert len(sd) == 50 and sd["dec_conv1a.weight"].shape == (96, 97, 3, 3) and all(v.dtype == torch.float32 for v in sd.values())
    assert all(torch.isfinite(v).all() for v in sd.values())
    assert not torch.equal(sd["nin_c.weight"], torch.load(ckpts[0], map_location="cpu")["nin_c.weight"])    # it trained

    _run([os.path.join(ROOT, "entry", "finetune.py"), "--synthetic", "2", "--pretrained_ckpt", ckpts[-1], "--n_epoch", "1",
          "--batchsize", "2", "--patch_size", "64", "--patches_per_image", "2", "--save_model_path", out, "--log_name", "ft"], ROOT)
    ad = torch.load(os.path.join(out, "ft", "epoch_adapter_001.pth"), map_location="cpu")
    assert len(ad) == 54 and "adapter.net.0.weight" in ad and "base.enc_conv0.weight" in ad

    ev = os.path.join(out, "eval")
    _run([os.path.join(ROOT, "entry", "evaluation.py"), "--synthetic", "2", "--checkpoint", ckpts[-1], "--save_dir", ev], ROOT)
    txt = open(os.path.join(ev, "metrics.txt")).read()
    assert txt.count("PSNR=") == 3 and "AVG" in txt
    _run([os.path.join(ROOT, "entry", "evaluation.py"), "--synthetic", "1", "--checkpoint", ckpts[-1], "--save_dir", ev, "--tiled"], ROOT)
    # evaluation_adapter.py: base + adapter from the finetune checkpoint
    ev2 = os.path.join(out, "eval_adapter")
    _run([os.path.join(ROOT, "entry", "evaluation.py"), "--synthetic", "1", "--adapter_ckpt", os.path.join(out, "ft", "epoch_adapter_001.pth"),
          "--save_dir", ev2], ROOT)
    assert "PSNR=" in open(os.path.join(ev2, "metrics.txt")).read()
    # per-snapshot validation + A_log.csv (train.py:391-434)
    logs = glob.glob(os.path.join(out, "UNET_test", "*", "validation", "A_log.csv"))
    assert len(logs) == 1 and open(logs[0]).read().count("epoch") == 2
    assert glob.glob(os.path.join(out, "UNET_test", "*", "validation", "*_denoised.png"))


def test_reference_named_eval_entry_points(tmp_path):
    """evaluation_704.py / evaluation_adapter.py under the reference's own names and flags (eval_704.sh:21-25,
    evaluation_adapter.py:17-44), and the fork's live supervised loop on RESNET (train.py:298-314, :354-368)."""
    import numpy as np
    from PIL import Image
    out = str(tmp_path)
    _run([os.path.join(ROOT, "entry", "train.py"), "--synthetic", "2", "--patch", "64", "--batchsize", "2", "--n_epoch", "1",
          "--loop", "supervised", "--save_model_path", out, "--log_name", "RESNET_sup", "--patches_per_image", "2"], ROOT)
    ck = sorted(glob.glob(os.path.join(out, "RESNET_sup", "*", "epoch_model_*.pth")))
    assert len(ck) == 2 and len(torch.load(ck[-1], map_location="cpu")) == 42
    _run([os.path.join(ROOT, "entry", "train.py"), "--synthetic", "2", "--patch", "64", "--batchsize", "2", "--n_epoch", "1",
          "--save_model_path", out, "--log_name", "UNET_n2n", "--patches_per_image", "2"], ROOT)
    ck = sorted(glob.glob(os.path.join(out, "UNET_n2n", "*", "epoch_model_*.pth")))
    ev = os.path.join(out, "eval704")
    txt = _run([os.path.join(ROOT, "entry", "evaluation_704.py"), "--synthetic", "2", "--checkpoint", ck[-1], "--save_dir", ev,
                "--log_name", "UNET_n2n"], ROOT)
    m = open(os.path.join(ev, "metrics.txt")).read()
    assert m.startswith("Average PSNR:") and "Average SSIM:" in m and "Average L1 Loss:" in m
    assert len(glob.glob(os.path.join(ev, "*_denoised.png"))) == 2
    d0 = np.array(Image.open(sorted(glob.glob(os.path.join(ev, "*_denoised.png")))[0]))
    assert (d0[0, :] == 0).all() and (d0[:, 0] == 0).all()          # evaluation_704.py's zero border weight
    # adapter inference: data_dir/noise (+ clean), --ckpt, --arch
    _run([os.path.join(ROOT, "entry", "finetune.py"), "--synthetic", "2", "--pretrained_ckpt", ck[-1], "--n_epoch", "1",
          "--batchsize", "2", "--patch_size", "64", "--patches_per_image", "2", "--save_model_path", out, "--log_name", "ft2"], ROOT)
    dd = os.path.join(out, "data")
    os.makedirs(os.path.join(dd, "noise")); os.makedirs(os.path.join(dd, "clean"))
    rng = np.random.RandomState(0)
    for i in range(2):
        c = rng.randint(0, 256, (64, 96)).astype(np.uint8)
        Image.fromarray(c).save(os.path.join(dd, "clean", f"im{i}.png"))
        Image.fromarray(np.clip(c + rng.randn(64, 96) * 25, 0, 255).astype(np.uint8)).save(os.path.join(dd, "noise", f"im{i}.png"))
    ev3 = os.path.join(out, "infer_adapter")
    txt = _run([os.path.join(ROOT, "entry", "evaluation_adapter.py"), "--data_dir", dd, "--ckpt", os.path.join(out, "ft2", "epoch_adapter_001.pth"),
                "--arch", "UNet", "--save_dir", ev3], ROOT)
    assert txt.count("PSNR=") == 2 and len(glob.glob(os.path.join(ev3, "*_denoised.png"))) == 2


def test_finetune_iqsl_entry_point(tmp_path):
    """finetune_iqsl.py under its own name and flags: IQSL term in the loss, adapter-only checkpoint (finetune_iqsl.py:114-132)."""
    out = str(tmp_path)
    _run([os.path.join(ROOT, "entry", "train.py"), "--synthetic", "2", "--patch", "64", "--batchsize", "2", "--n_epoch", "1",
          "--save_model_path", out, "--log_name", "UNET_b", "--patches_per_image", "2"], ROOT)
    ck = sorted(glob.glob(os.path.join(out, "UNET_b", "*", "epoch_model_*.pth")))
    txt = _run([os.path.join(ROOT, "entry", "finetune_iqsl.py"), "--synthetic", "2", "--arch", "UNet", "--pretrained_ckpt", ck[-1],
                "--n_epoch", "1", "--batchsize", "2", "--patch_size", "64", "--patches_per_image", "2", "--save_model_path", out,
                "--log_name", "ftq", "--lambda_iqsl", "0.2", "--iqsl_margin", "0.01"], ROOT)
    assert "[IQSL] Estimated thresholds" in txt and "IQSL=" in txt
    ad = torch.load(os.path.join(out, "ftq", "epoch_adapter_only_001.pth"), map_location="cpu")
    assert sorted(ad.keys()) == ["net.0.bias", "net.0.weight", "net.2.bias", "net.2.weight"]
    assert all(torch.isfinite(v).all() for v in ad.values())


def test_improved_unet_through_the_entry_points(tmp_path):
    """train.sh:3 / eval_704.sh launch the 'UNetImproved' family (arch_unet.ImprovedUNet): the fork's live supervised loop,
    then evaluation_704.py on its checkpoint (reference state_dict keys)."""
    out = str(tmp_path)
    _run([os.path.join(ROOT, "entry", "train.py"), "--synthetic", "2", "--patch", "64", "--batchsize", "2", "--n_epoch", "1",
          "--loop", "supervised", "--save_model_path", out, "--log_name", "UNetImproved_t", "--patches_per_image", "2",
          "--n_feature", "16"], ROOT)
    ck = sorted(glob.glob(os.path.join(out, "UNetImproved_t", "*", "epoch_model_*.pth")))
    sd = torch.load(ck[-1], map_location="cpu")
    assert "ups.3.res.block.4.bias" in sd and "noise_estimator.2.weight" in sd and sd["final.weight"].shape == (1, 9, 3, 3)
    ev = os.path.join(out, "eval704")
    _run([os.path.join(ROOT, "entry", "evaluation_704.py"), "--synthetic", "1", "--checkpoint", ck[-1], "--save_dir", ev,
          "--log_name", "UNetImproved_t", "--n_feature", "16"], ROOT)
    assert "Average PSNR:" in open(os.path.join(ev, "metrics.txt")).read()


def test_n2n_loop_on_other_network_families(tmp_path):
    """--loop n2n with RESNET / UNetImproved: the iteration of training_script.md:137-156 through the drop-in functions and
    autograd (the fused N2NTrainer is built for arch_unet.UNet)."""
    out = str(tmp_path)
    for log_name, extra, nkeys in (("RESNET_n2n", [], 42), ("UNetImproved_n2n", ["--n_feature", "16"], None)):
        txt = _run([os.path.join(ROOT, "entry", "train.py"), "--synthetic", "2", "--patch", "64", "--batchsize", "2", "--n_epoch", "1",
                    "--save_model_path", out, "--log_name", log_name, "--patches_per_image", "2"] + extra, ROOT)
        assert "Loss_Full=" in txt and "Lambda=" in txt
        ck = sorted(glob.glob(os.path.join(out, log_name, "*", "epoch_model_*.pth")))
        a, b = torch.load(ck[0], map_location="cpu"), torch.load(ck[-1], map_location="cpu")
        assert len(ck) == 2 and (nkeys is None or len(b) == nkeys)
        assert all(torch.isfinite(v).all() for v in b.values()) and any(not torch.equal(a[k], b[k]) for k in a)
