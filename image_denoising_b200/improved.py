"""Drop-in ``arch_unet.ImprovedUNet`` (reference arch_unet.py:420-531).

Same constructor, sub-module tree and therefore the same ``state_dict`` keys / shapes / default initialisation stream as
the reference (``noise_estimator.{0,2}``, ``downs.{i}.{0,2,3}``, ``bottle.{0,1}``, ``ups.{i}.{conv_ps,fuse,rdb,res}``,
``final``); the ``nn`` sub-modules only HOLD the parameters.  Every arithmetic operator of the forward and backward pass is
a kernel of libn2n_b200 reached through the C-ABI (include/n2n_b200.h):

* 3x3 / 1x1 convolutions (forward, input gradient, weight + bias gradient): ``n2n_conv2d_{fwd,dgrad,wgrad}`` — the tcgen05
  engines in bf16, the CUDA-core parity engine in fp32.  Layers wider than one launch of the engine allows (256 accumulator
  columns, 128 x 144 channels per weight-gradient launch) are issued as channel chunks;
* GroupNorm (+ fused LeakyReLU or residual add), LeakyReLU / Sigmoid, PixelShuffle(2), MaxPool2d(2), residual adds:
  ``n2n_groupnorm_{fwd,bwd}``, ``n2n_act_{fwd,bwd}``, ``n2n_pixel_shuffle2``, ``n2n_maxpool2_{fwd,bwd}``, ``n2n_add_f32``.

Forward and backward run on the native executor (csrc/improved_plan.cu: ``n2n_improved_forward`` / ``n2n_improved_backward``,
one C call each, activations and gradients resident in the engines' blocked layout, dense / skip concats written in place).
The per-layer composition above — PyTorch's tape over single-layer C-ABI calls, fp32 NCHW between layers — is kept as the
cross-check of the executor (``net.native_train = False``) and serves inputs that require grad."""
from __future__ import annotations

import ctypes
import os

import torch
import torch.nn as nn


from . import _ext, ops
from ._ext import check, lib, ptr, ptr_array, require_cuda, stream_ptr


_DEBUG_TAPE = None      # tests: a list that collects (name, tensor) of selected intermediates of the per-layer path
_NATIVE_NOGRAD = os.environ.get("N2N_IMPROVED_NATIVE", "1") != "0"      # 0: no-grad calls also take the layer-by-layer path
_NATIVE_TRAIN = os.environ.get("N2N_IMPROVED_NATIVE_TRAIN", "1") != "0"  # 0: training composes the per-layer calls under autograd


class _ImprovedFunction(torch.autograd.Function):
    """Forward + backward of the whole network on the native executor (csrc/improved_plan.cu); parameter gradients only
    (a network input that requires grad takes the layer-by-layer path)."""

    @staticmethod
    def forward(ctx, net, x, *params):
        key, plan, slot = net._checkout(x, True)
        net._param_list(plan)
        slot["x"].copy_(x)
        check(lib().n2n_improved_forward(plan, ptr_array(params), ptr(slot["x"]), ptr(slot["y"]), ptr(slot["ws"]), stream_ptr()))
        net.last_launches = lib().n2n_improved_launches(plan, 0)
        ctx.net, ctx.key, ctx.plan, ctx.slot, ctx.params = net, key, plan, slot, params
        return slot["y"].clone()            # slot["y"] itself stays put for the backward (sigmoid')

    @staticmethod
    def backward(ctx, dy):
        slot = ctx.slot
        slot["dy"].copy_(dy)

        def views_of(flat):
            out, off = [], 0
            for q, nq in zip(ctx.params, slot["sizes"]):
                out.append(flat[off:off + nq].view(q.shape))
                off += nq
            return out

        check(lib().n2n_improved_backward(ctx.plan, ptr_array(ctx.params), ptr(slot["dy"]), ptr(slot["y"]), ptr_array(views_of(slot["g"])),
                                          ptr(slot["ws"]), stream_ptr()))
        ctx.net.last_bwd_launches = lib().n2n_improved_launches(ctx.plan, 1)
        grads = views_of(slot["g"].clone())                  # one copy; the gradients handed to autograd are views of it
        ctx.net._last_train = (ctx.plan, slot["ws"])         # read_buffer() hook: valid until the workspace is reused
        ctx.net._give_back(ctx.key, slot)
        ctx.slot = None
        return (None, None) + tuple(grads)


# ----------------------------------------------------------------------------- channel chunking of wide layers
def _chunks(total: int, limit: int):
    n = -(-total // limit)
    size = -(-total // n)
    size = -(-size // 16) * 16 if n > 1 else size
    return [(c0, min(c0 + size, total)) for c0 in range(0, total, size)]


def _conv_fwd(x, w, b, slope, precision):
    cout = w.shape[0]
    if precision != "bf16" or cout <= 256:
        return ops.conv2d_fwd(x, w, b, act_slope=slope, precision=precision)
    parts = [ops.conv2d_fwd(x, w[c0:c1], None if b is None else b[c0:c1], act_slope=slope, precision=precision)
             for c0, c1 in _chunks(cout, 256)]
    return torch.cat(parts, dim=1)


def _conv_dgrad(dy, w, precision):
    cin = w.shape[1]
    if precision != "bf16" or cin <= 256:
        return ops.conv2d_dgrad(dy, w, precision=precision)
    return torch.cat([ops.conv2d_dgrad(dy, w[:, c0:c1], precision=precision) for c0, c1 in _chunks(cin, 256)], dim=1)


def _conv_wgrad(x, dy, k, precision):
    cin, cout = x.shape[1], dy.shape[1]
    if precision != "bf16" or (cin <= 144 and cout <= 128):
        return ops.conv2d_wgrad(x, dy, k, precision=precision)
    dw = torch.empty((cout, cin, k, k), dtype=torch.float32, device=x.device)
    db = torch.empty((cout,), dtype=torch.float32, device=x.device)
    xs = [(c0, c1, x[:, c0:c1].contiguous()) for c0, c1 in _chunks(cin, 144)]
    for o0, o1 in _chunks(cout, 128):
        dyo = dy[:, o0:o1].contiguous()
        for c0, c1, xc in xs:
            dwp, dbp = ops.conv2d_wgrad(xc, dyo, k, precision=precision)
            dw[o0:o1, c0:c1] = dwp
        db[o0:o1] = dbp
    return dw, db


# ----------------------------------------------------------------------------- autograd nodes over the C-ABI kernels
class _Conv(torch.autograd.Function):
    """Conv2d(k in {1,3}, stride 1, 'same') with an optional fused LeakyReLU (slope >= 0)."""

    @staticmethod
    def forward(ctx, x, w, b, slope, precision):
        y = _conv_fwd(x, w, b, slope, precision)
        ctx.slope, ctx.precision, ctx.has_bias = slope, precision, b is not None
        ctx.save_for_backward(x, w, y if slope >= 0 else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.slope >= 0:
            dy = ops.act_bwd(y, dy, ops.ACT_LRELU, ctx.slope)
        dx = _conv_dgrad(dy, w, ctx.precision) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = _conv_wgrad(x, dy, w.shape[2], ctx.precision)
        return dx, dw, (db if ctx.has_bias else None), None, None


class _GroupNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, slope, residual):
        y, stats = ops.groupnorm_fwd(x, gamma, beta, groups, eps, slope, residual, want_stats=True)
        ctx.groups, ctx.slope, ctx.has_res = groups, slope, residual is not None
        ctx.save_for_backward(x, gamma, y if slope >= 0 else None, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, y, stats = ctx.saved_tensors
        dy = dy.contiguous()
        dx, dgamma, dbeta = ops.groupnorm_bwd(x, gamma, y, dy, stats, ctx.groups, ctx.slope)
        return dx, dgamma, dbeta, None, None, None, (dy if ctx.has_res else None)


class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind, slope):
        y = ops.act_fwd(x, kind, slope)
        ctx.kind, ctx.slope = kind, slope
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.act_bwd(y, dy.contiguous(), ctx.kind, ctx.slope), None, None


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


class _Pool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.maxpool2_fwd(x, precision="fp32")

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.maxpool2_bwd(x, dy.contiguous(), 1.0, precision="fp32")      # gradient to the first maximum (ATen's rule)


class _PixelShuffle(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.pixel_shuffle2(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.pixel_shuffle2(dy.contiguous(), inverse=True)


def _conv(x, conv: nn.Conv2d, precision: str, slope: float = -1.0):
    return _Conv.apply(x, conv.weight, conv.bias, slope, precision)


def _gn(x, gn: nn.GroupNorm, slope: float = -1.0, residual=None):
    return _GroupNorm.apply(x, gn.weight, gn.bias, gn.num_groups, gn.eps, slope, residual)


# ----------------------------------------------------------------------------- parameter holders (reference module tree)
def _norm_gn(channels: int, groups: int = 32) -> nn.GroupNorm:
    """arch_unet.py:7-15, kind 'gn'."""
    g = min(groups, channels)
    while channels % g != 0 and g > 1:
        g -= 1
    return nn.GroupNorm(g, channels, affine=True)


class ResBlock(nn.Module):
    """arch_unet.py:420-432: x + GN(conv(LReLU(GN(conv(x))))), convolutions without bias."""

    def __init__(self, channels):
        super().__init__()
        self.block = nn.Sequential(
            nn.Conv2d(channels, channels, 3, 1, 1, bias=False), _norm_gn(channels),
            nn.LeakyReLU(0.2, True),
            nn.Conv2d(channels, channels, 3, 1, 1, bias=False), _norm_gn(channels))

    def run(self, x, precision):
        b = self.block
        t1 = _conv(x, b[0], precision)
        t = _gn(t1, b[1], slope=0.2)
        t3 = _conv(t, b[3], precision)
        if _DEBUG_TAPE is not None and t.requires_grad:
            for q in (t1, t, t3):
                q.retain_grad()
            _DEBUG_TAPE.append((t1, t, t3))
        return _gn(t3, b[4], residual=x)


class RDB(nn.Module):
    """arch_unet.py:435-449: four dense 3x3 convolutions (growth 32, LeakyReLU) + 1x1 local feature fusion + residual."""

    def __init__(self, channels, growth=32, layers=4):
        super().__init__()
        self.convs = nn.ModuleList()
        in_ch = channels
        for _ in range(layers):
            self.convs.append(nn.Conv2d(in_ch, growth, 3, 1, 1, bias=True))
            in_ch += growth
        self.lff = nn.Conv2d(in_ch, channels, 1, 1, 0, bias=True)
        self.act = nn.LeakyReLU(0.2, True)

    def run(self, x, precision):
        feats = x
        for conv in self.convs:
            feats = torch.cat([feats, _conv(feats, conv, precision, slope=0.2)], dim=1)
        return _Add.apply(x, _conv(feats, self.lff, precision))


class UpBlock(nn.Module):
    """arch_unet.py:452-472: conv3x3 -> PixelShuffle(2) -> cat skip -> fuse conv3x3 + LeakyReLU -> RDB -> ResBlock."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv_ps = nn.Conv2d(in_ch, out_ch * 4, 3, 1, 1, bias=True)
        self.ps = nn.PixelShuffle(2)
        self.fuse = nn.Conv2d(out_ch * 3, out_ch, 3, 1, 1, bias=True)
        self.rdb = RDB(out_ch)
        self.res = ResBlock(out_ch)

    def run(self, x, skip, precision):
        x = _PixelShuffle.apply(_conv(x, self.conv_ps, precision))
        x = _conv(torch.cat([x, skip], dim=1), self.fuse, precision, slope=0.2)
        return self.res.run(self.rdb.run(x, precision), precision)


class ImprovedUNet(nn.Module):
    """arch_unet.py:475-531.  ``forward`` needs H and W divisible by 2**depth (the reference fails in ``torch.cat`` otherwise)."""

    def __init__(self, in_nc=3, out_nc=3, n_feature=48, depth=4, noise=True):
        super().__init__()
        self.in_nc = in_nc
        self.out_nc = out_nc
        self.n_feature = n_feature
        self.depth = depth
        self.noise = noise
        if self.noise:
            self.noise_estimator = nn.Sequential(
                nn.Conv2d(in_nc, n_feature, 3, 1, 1, bias=True), nn.LeakyReLU(0.2, True),
                nn.Conv2d(n_feature, 1, 3, 1, 1, bias=True), nn.Sigmoid())
        self.downs, self.pools = nn.ModuleList(), nn.ModuleList()
        nf = n_feature
        for i in range(depth):
            inc = (in_nc + 1 if self.noise else 1) if i == 0 else nf // 2          # arch_unet.py:493-498 (noise=False assumes in_nc == 1)
            self.downs.append(nn.Sequential(nn.Conv2d(inc, nf, 3, 1, 1, bias=True), nn.LeakyReLU(0.2, True), RDB(nf), ResBlock(nf)))
            self.pools.append(nn.MaxPool2d(2))
            nf *= 2
        self.bottle = nn.Sequential(RDB(nf // 2), ResBlock(nf // 2))
        nf = nf // 2
        self.ups = nn.ModuleList()
        for _ in range(depth):
            self.ups.append(UpBlock(nf, nf // 2))
            nf //= 2
        self.final = nn.Conv2d(n_feature // 2 + in_nc, out_nc, 3, 1, 1, bias=True)
        self.sigmoid = nn.Sigmoid()
        self.precision = _ext.default_precision()
        # native executor (csrc/improved_plan.cu) for no-grad calls / for forward + backward; False = compose the per-layer
        # C-ABI calls under autograd (kept as the cross-check of the executor and for inputs that require grad)
        self.native_nograd = _NATIVE_NOGRAD
        self.native_train = _NATIVE_TRAIN

    def set_precision(self, precision: str):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        return self

    def _checkout(self, x, train):
        """(key, plan, workspace) for this shape / precision; training workspaces are held until the backward has run, so two
        forwards before one backward (train.py:361) keep separate activations."""
        n, _, h, w = x.shape
        key = (n, h, w, _ext.dtype_tag(self.precision), x.device.index, bool(train))
        plans = self.__dict__.setdefault("_plans", {})
        pools = self.__dict__.setdefault("_free_ws", {})
        if key not in plans:
            if len(plans) >= 6:
                old_key = next(iter(plans))
                lib().n2n_improved_plan_destroy(plans.pop(old_key))
                pools.pop(old_key, None)
            handle = ctypes.c_void_p()
            check(lib().n2n_improved_plan_create(ctypes.byref(handle), self.in_nc, self.out_nc, self.n_feature, self.depth,
                                                 int(self.noise), n, h, w, key[3], int(train)))
            plans[key] = handle
        plan = plans[key]
        pool = pools.setdefault(key, [])
        if pool:
            return key, plan, pool.pop()
        # workspace + the call's boundary tensors kept IN PLACE (input, output, dL/dy, flat parameter gradients): the executor
        # caches its launch sequence as a CUDA graph keyed by these pointers
        f32 = dict(dtype=torch.float32, device=x.device)
        slot = {"ws": torch.empty(lib().n2n_improved_workspace_bytes(plan), dtype=torch.uint8, device=x.device),
                "x": torch.empty((n, self.in_nc, h, w), **f32), "y": torch.empty((n, self.out_nc, h, w), **f32)}
        if train:
            slot["dy"] = torch.empty((n, self.out_nc, h, w), **f32)
            slot["sizes"] = [q.numel() for q in self.parameters()]
            slot["g"] = torch.empty(sum(slot["sizes"]), **f32)
        return key, plan, slot

    def _give_back(self, key, slot):
        pool = self.__dict__.setdefault("_free_ws", {}).setdefault(key, [])
        if len(pool) < 2:
            pool.append(slot)

    def _param_list(self, plan):
        params = list(self.parameters())
        if len(params) != lib().n2n_improved_num_params(plan):
            raise RuntimeError("ImprovedUNet: parameter list does not match the native plan")
        for q in params:
            require_cuda(q, "ImprovedUNet parameters")
            if q.dtype != torch.float32 or not q.is_contiguous():
                raise RuntimeError("ImprovedUNet parameters must be contiguous float32 tensors")
        return params

    def _native_forward(self, x):
        """n2n_improved_forward on a (plan, workspace) cached per shape / precision."""
        key, plan, slot = self._checkout(x, False)
        params = self._param_list(plan)
        slot["x"].copy_(x)
        check(lib().n2n_improved_forward(plan, ptr_array(params), ptr(slot["x"]), ptr(slot["y"]), ptr(slot["ws"]), stream_ptr()))
        self.last_launches = lib().n2n_improved_launches(plan, 0)
        y = slot["y"].clone()
        self._give_back(key, slot)
        return y

    def __getstate__(self):
        # native plans / workspaces are per-process handles: copies and pickles of the module start without them
        state = self.__dict__.copy()
        for k in ("_plans", "_free_ws", "_last_train"):
            state.pop(k, None)
        return state

    def read_buffer(self, buf: int, grad: bool = False):
        """Layer-level parity hook: buffer ``buf`` of the last native training step (its gradient mirror with grad=True) as
        fp32 NCHW, all 16-channel blocks (zero padding included)."""
        plan, ws = self._last_train
        dims = (ctypes.c_int * 3)()
        n = lib().n2n_improved_read_buffer(plan, ptr(ws), buf, int(grad), None, dims, stream_ptr())
        if n <= 0:
            raise ValueError("no such buffer")
        out = torch.empty((n // (dims[0] * dims[1] * dims[2]), dims[0], dims[1], dims[2]), dtype=torch.float32, device=ws.device)
        if lib().n2n_improved_read_buffer(plan, ptr(ws), buf, int(grad), ptr(out), dims, stream_ptr()) < 0:
            raise RuntimeError("read_buffer failed")
        return out

    def __del__(self):
        try:
            for plan in self.__dict__.get("_plans", {}).values():
                lib().n2n_improved_plan_destroy(plan)
        except Exception:
            pass

    def forward(self, x):
        require_cuda(x, "ImprovedUNet.forward")
        if x.dim() != 4 or x.shape[1] != self.in_nc:
            raise ValueError(f"expected input [N,{self.in_nc},H,W], got {tuple(x.shape)}")
        m = 1 << self.depth
        if x.shape[2] % m or x.shape[3] % m:
            raise ValueError(f"H and W must be multiples of {m} ({self.depth} 2x2 poolings, arch_unet.py:521-523)")
        p = self.precision
        x = x.contiguous().float()
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(q.requires_grad for q in self.parameters()))
        if self.native_nograd and not need_grad:
            return self._native_forward(x)
        if self.native_train and need_grad and not x.requires_grad:
            return _ImprovedFunction.apply(self, x, *self.parameters())
        if self.noise:
            ne = self.noise_estimator
            sigma = _Act.apply(_conv(_conv(x, ne[0], p, slope=0.2), ne[2], p), ops.ACT_SIGMOID, 0.0)
            x = torch.cat([x, sigma], dim=1)
        orig = x[:, :self.in_nc]
        skips = []
        for down in self.downs:
            x = down[3].run(down[2].run(_conv(x, down[0], p, slope=0.2), p), p)
            skips.append(x)
            x = _Pool.apply(x)
        x = self.bottle[1].run(self.bottle[0].run(x, p), p)
        for up, skip in zip(self.ups, reversed(skips)):
            x = up.run(x, skip, p)
        x = _conv(torch.cat([x, orig], dim=1), self.final, p)
        return _Act.apply(x, ops.ACT_SIGMOID, 0.0)
