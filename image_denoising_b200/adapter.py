"""Drop-in ``adapter.OutputAdapter`` / ``adapter.DenoiserWithAdapter`` (reference adapter.py:5-67).

State-dict layout is the reference's 54 keys: ``base.*`` (50) + ``adapter.net.0.weight``
[16,2C,3,3], ``adapter.net.0.bias``, ``adapter.net.2.weight`` [C,16,3,3], ``adapter.net.2.bias``.
The two adapter convolutions, the ReLU, the concat and the residual add run as
n2n_adapter_forward / n2n_adapter_backward (same engines as the UNet)."""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _ext
from ._ext import check, lib, ptr, ptr_array, require_cuda, stream_ptr


class _AdapterFunction(torch.autograd.Function):
    """The workspace (cat buffer + hidden activations) is checked out of the module's pool until backward has
    run, so two forwards before one backward (gradient accumulation) keep separate activations."""

    @staticmethod
    def forward(ctx, mod, noisy, base_out, *params):
        key, plan, ws = mod._checkout(noisy, True)
        out = torch.empty_like(base_out)
        check(lib().n2n_adapter_forward(plan, ptr_array(params), ptr(noisy), ptr(base_out), ptr(out), ptr(ws), stream_ptr()))
        ctx.mod, ctx.key, ctx.plan, ctx.ws, ctx.params = mod, key, plan, ws, params
        ctx.inputs = (noisy, base_out)           # conv1's weight gradient reads the concat operand again
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous().float()
        grads = [torch.empty_like(p) for p in ctx.params]
        check(lib().n2n_adapter_backward(ctx.plan, ptr_array(ctx.params), ptr(ctx.inputs[0]), ptr(ctx.inputs[1]), ptr(dout),
                                         ptr_array(grads), ptr(ctx.ws), stream_ptr()))
        ctx.mod._give_back(ctx.key, ctx.ws)
        ctx.ws = ctx.inputs = None
        return (None, None, None) + tuple(grads)


class OutputAdapter(nn.Module):
    def __init__(self, in_channels: int = 1, hidden_channels: int = 16):
        super().__init__()
        self.in_channels = in_channels
        self.hidden_channels = hidden_channels
        self.net = nn.Sequential(      # parameter holders; default torch Conv2d init as in the reference
            nn.Conv2d(2 * in_channels, hidden_channels, kernel_size=3, padding=1, bias=True),
            nn.ReLU(inplace=True),
            nn.Conv2d(hidden_channels, in_channels, kernel_size=3, padding=1, bias=True),
        )
        self.precision = _ext.default_precision()
        self._plans = {}       # key -> plan handle
        self._free_ws = {}     # key -> workspaces not held by an autograd graph

    def _checkout(self, x, bwd):
        n, c, h, w = x.shape
        key = (n, h, w, _ext.dtype_tag(self.precision), bool(bwd), x.device.index)
        if key not in self._plans:
            handle = ctypes.c_void_p()
            check(lib().n2n_adapter_plan_create(ctypes.byref(handle), self.in_channels, self.hidden_channels,
                                                n, h, w, key[3], int(bwd)))
            self._plans[key] = handle
        pool = self._free_ws.setdefault(key, [])
        ws = pool.pop() if pool else torch.empty(lib().n2n_adapter_workspace_bytes(self._plans[key]),
                                                 dtype=torch.uint8, device=x.device)
        return key, self._plans[key], ws

    def _give_back(self, key, ws):
        pool = self._free_ws.setdefault(key, [])
        if len(pool) < 2:
            pool.append(ws)

    def __getstate__(self):
        # plan handles / workspaces are per-process native objects: copies and pickles of the module start without them
        state = self.__dict__.copy()
        state["_plans"], state["_free_ws"] = {}, {}
        return state

    def clear_plans(self):
        for h in self._plans.values():
            lib().n2n_adapter_plan_destroy(h)
        self._plans.clear()
        self._free_ws.clear()

    def __del__(self):
        try:
            self.clear_plans()
        except Exception:
            pass

    def forward(self, noisy: torch.Tensor, base_out: torch.Tensor) -> torch.Tensor:
        require_cuda(noisy, "OutputAdapter.forward")
        noisy = noisy.contiguous().float()
        base_out = base_out.contiguous().float()
        params = [self.net[0].weight, self.net[0].bias, self.net[2].weight, self.net[2].bias]
        if torch.is_grad_enabled() and (base_out.requires_grad or noisy.requires_grad):
            # adapter.py:22-26 lets gradients flow into base_out (d out / d base_out = I + the conv path); the
            # B200 path implements the frozen-base finetune of finetune.py:238-263 only: fail loudly instead
            # of silently detaching the base network.
            raise NotImplementedError(
                "OutputAdapter: gradients with respect to noisy / base_out are not implemented on the B200 path "
                "(use freeze_base=True, use_no_grad_for_base=True as finetune.py:238-246 does)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _AdapterFunction.apply(self, noisy, base_out, *params)
        key, plan, ws = self._checkout(noisy, False)
        out = torch.empty_like(base_out)
        check(lib().n2n_adapter_forward(plan, ptr_array(params), ptr(noisy), ptr(base_out), ptr(out), ptr(ws), stream_ptr()))
        self._give_back(key, ws)
        return out


class DenoiserWithAdapter(nn.Module):
    def __init__(self, base_model: nn.Module, in_channels: int = 1, hidden_channels: int = 16,
                 freeze_base: bool = True, use_no_grad_for_base: bool = True):
        super().__init__()
        self.base = base_model
        self.in_channels = in_channels
        self.freeze_base = freeze_base
        self.use_no_grad_for_base = use_no_grad_for_base
        if freeze_base:
            for p in self.base.parameters():
                p.requires_grad = False
        self.adapter = OutputAdapter(in_channels=in_channels, hidden_channels=hidden_channels)

    def set_precision(self, precision: str):
        self.adapter.precision = precision
        if hasattr(self.base, "set_precision"):
            self.base.set_precision(precision)
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.use_no_grad_for_base:
            with torch.no_grad():
                base_out = self.base(x)
        else:
            base_out = self.base(x)
        return self.adapter(x, base_out)
