"""FusedAdam — torch.optim.Adam semantics (defaults of train.py:332 / finetune.py:260-263:
betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) as ONE multi-tensor kernel launch
(n2n_adam_multi) per param group.  Subclasses torch.optim.Optimizer so lr schedulers
(MultiStepLR, train.py:333-340) and state_dict() work unchanged; state keys match torch's
('step', 'exp_avg', 'exp_avg_sq')."""
from __future__ import annotations

import torch

from . import _ext
from ._ext import check, lib, ptr, stream_ptr


def multistep_milestones(n_epoch: int):
    """train.py:333-340: MultiStepLR milestones int(20r)-1, int(40r)-1, int(60r)-1, int(80r)-1, r = n_epoch/100."""
    ratio = n_epoch / 100
    return [int(20 * ratio) - 1, int(40 * ratio) - 1, int(60 * ratio) - 1, int(80 * ratio) - 1]


def multistep_lr(base_lr: float, epoch: int, n_epoch: int, gamma: float) -> float:
    """Learning rate in force during the 1-based ``epoch`` of the reference's loop (train.py:333-340, :346,
    :375): ``scheduler.step()`` runs at the END of every epoch, so epoch e trains with last_epoch = e - 1 and
    torch's MultiStepLR has decayed once for every milestone m <= e - 1."""
    return base_lr * gamma ** sum(1 for m in multistep_milestones(n_epoch) if (epoch - 1) >= m)


def build_adam_tables(params, grads, exp_avg, exp_avg_sq, device):
    """Device tables for n2n_adam_multi: int64 [ntensors,5] and int32 [nblocks,2]."""
    rows, blocks = [], []
    for t, (p, g, m, v) in enumerate(zip(params, grads, exp_avg, exp_avg_sq)):
        n = p.numel()
        rows.append([p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n])
        for c in range((n + _ext.ADAM_CHUNK - 1) // _ext.ADAM_CHUNK):
            blocks.append([t, c])
    table = torch.tensor(rows, dtype=torch.int64).to(device)
    blk = torch.tensor(blocks, dtype=torch.int32).to(device)
    return table, blk


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.grad_scale = grad_scale
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                _ext.require_cuda(p, "FusedAdam")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous float32 parameters")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
            sig = tuple((p.data_ptr(), g.data_ptr(), self.state[p]["exp_avg"].data_ptr()) for p, g in zip(ps, grads))
            cached = self._tables.get(gi)
            if cached is None or cached[0] != sig:
                table, blk = build_adam_tables(ps, grads, [self.state[p]["exp_avg"] for p in ps],
                                               [self.state[p]["exp_avg_sq"] for p in ps], ps[0].device)
                cached = (sig, table, blk)
                self._tables[gi] = cached
            _, table, blk = cached
            step = int(self.state[ps[0]]["step"]) + 1
            for p in ps:
                self.state[p]["step"] = step
            b1, b2 = group["betas"]
            check(lib().n2n_adam_multi(ptr(table), len(ps), ptr(blk), blk.shape[0], float(group["lr"]), float(b1),
                                       float(b2), float(group["eps"]), step, float(self.grad_scale), stream_ptr()))
        return loss
