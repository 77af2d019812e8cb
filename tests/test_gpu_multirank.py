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
