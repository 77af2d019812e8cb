"""On-hardware multi-rank parity (SURVEY.md §4 tier 4): launched under torchrun on 2 GPUs (skipped on a
1-GPU box): W ranks x B patches reproduce the 1-process W*B-batch gradient (same masks by construction:
the selector is drawn for the global batch and sliced), and all replicas hold bit-identical weights after
five optimiser steps."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_two_rank_nccl_gradient_equals_global_batch():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "dp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DP_PARITY ")][-1]
    res = json.loads(line[len("DP_PARITY "):])
    assert res["world"] == 2 and res["replicas_identical"] is True, res
    assert res["fp32"]["ok"] and res["fp32"]["max_rel"] <= 2e-4, res      # fp32 engine: summation order only
    assert res["bf16"]["ok"], res


@pytest.mark.timeout(400)
def test_two_rank_entry_train_runs_and_exits(tmp_path):
    """entry/train.py --parallel under torchrun (what replaces nn.DataParallel, train.py:324-325): the fused N2N loop and the
    autograd supervised loop train one epoch on 2 ranks, rank 0 writes the reference's checkpoints, and the processes EXIT
    (the captured step graph holds NCCL work and must be dropped before the process group is destroyed)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import glob
    for log_name, extra in (("UNET_dp", ["--loop", "n2n"]), ("UNetImproved_dp", ["--loop", "supervised", "--n_feature", "16"])):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(ROOT, "entry", "train.py"), "--synthetic", "4", "--patch", "64", "--batchsize", "4",
               "--n_epoch", "1", "--parallel", "--patches_per_image", "4", "--save_model_path", str(tmp_path), "--log_name", log_name] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=180, cwd=ROOT)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        assert "batch 2/GPU" in r.stdout and len(glob.glob(os.path.join(str(tmp_path), log_name, "*", "epoch_model_001.pth"))) == 1
