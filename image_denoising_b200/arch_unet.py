"""Drop-in ``arch_unet.UNet`` (reference arch_unet.py:100-260, non-blindspot branch).

Same constructor signature, the same 50 parameters under the same names / shapes /
registration order (so ``state_dict()``, ``load_state_dict()`` and
``optim.Adam(network.parameters())`` behave as with the reference), the same
initialisation (kaiming_normal fan_in x 0.1, zero bias — arch_unet.py:24-48, drawn in
the same order from torch's global RNG), but ``forward`` runs the whole network
through libn2n_b200's native executor (n2n_unet_forward / n2n_unet_backward) instead
of nn.Conv2d arithmetic.  There is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.init as init

from . import _ext
from ._ext import check, lib, ptr, ptr_array, require_cuda, stream_ptr


def initialize_weights(net_l, scale=1):
    """arch_unet.py:24-48 (conv / deconv / linear / BN branches)."""
    if not isinstance(net_l, list):
        net_l = [net_l]
    for net in net_l:
        for m in net.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d, nn.Linear)):
                init.kaiming_normal_(m.weight, a=0, mode='fan_in')
                m.weight.data *= scale
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
                init.constant_(m.weight, 1)
                init.constant_(m.bias.data, 0.0)


class UpsampleCat(nn.Module):
    """Parameter holder for arch_unet.py:51-62 (ConvTranspose2d k2 s2 + concat)."""

    def __init__(self, in_nc, out_nc):
        super().__init__()
        self.in_nc = in_nc
        self.out_nc = out_nc
        self.deconv = nn.ConvTranspose2d(in_nc, out_nc, 2, 2, 0, 0)
        initialize_weights(self.deconv, 0.1)

    def forward(self, x1, x2):  # pragma: no cover - the fused executor never calls this
        raise RuntimeError("UpsampleCat is executed inside UNet.forward by the native engine")


class _PlanCache:
    """(plan, workspace) pairs keyed by shape; workspaces used by an autograd graph are
    checked out until the backward has run."""

    def __init__(self):
        self.plans = {}
        self.free_ws = {}

    def plan(self, key):
        if key not in self.plans:
            arch, in_nc, out_nc, nf, n, h, w, dt, bwd = key
            handle = _ext.c_void_p()
            create = lib().n2n_resnet_plan_create if arch == "resnet" else lib().n2n_unet_plan_create
            check(create(_ext.ctypes.byref(handle), in_nc, out_nc, nf, n, h, w, dt, int(bwd)))
            self.plans[key] = handle
        return self.plans[key]

    def checkout(self, key, device):
        plan = self.plan(key)
        pool = self.free_ws.setdefault((key, device), [])
        if pool:
            return plan, pool.pop()
        nbytes = lib().n2n_unet_workspace_bytes(plan)
        return plan, torch.empty(nbytes, dtype=torch.uint8, device=device)

    def give_back(self, key, device, ws):
        pool = self.free_ws.setdefault((key, device), [])
        if len(pool) < 2:
            pool.append(ws)

    def clear(self):
        for h in self.plans.values():
            lib().n2n_unet_plan_destroy(h)
        self.plans.clear()
        self.free_ws.clear()

    # plan handles are per-process native objects: a copied / unpickled module starts with an empty cache
    def __deepcopy__(self, memo):
        return _PlanCache()

    def __reduce__(self):
        return (_PlanCache, ())


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, *params):
        key = net._plan_key(x, True)
        plan, ws = net._cache.checkout(key, x.device)
        y = torch.empty((x.shape[0], net.out_nc, x.shape[2], x.shape[3]), dtype=torch.float32, device=x.device)
        check(lib().n2n_unet_forward(plan, ptr_array(params), ptr(x), ptr(y), ptr(ws), stream_ptr()))
        ctx.net, ctx.key, ctx.plan, ctx.ws = net, key, plan, ws
        ctx.params = params
        ctx.need_dx = x.requires_grad
        ctx.xshape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        net = ctx.net
        dy = dy.contiguous().float()
        sizes = [p.numel() for p in ctx.params]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=dy.device)
        grads = [g.view(p.shape) for g, p in zip(flat.split(sizes), ctx.params)]
        dx = torch.empty(ctx.xshape, dtype=torch.float32, device=dy.device) if ctx.need_dx else None
        check(lib().n2n_unet_backward(ctx.plan, ptr_array(ctx.params), ptr(dy), ptr_array(grads), ptr(dx),
                                      ptr(ctx.ws), stream_ptr()))
        net._cache.give_back(ctx.key, dy.device, ctx.ws)
        ctx.ws = None
        unused = getattr(net, "_unused_params", ())
        return (None, dx) + tuple(None if i in unused else g for i, g in enumerate(grads))


class UNet(nn.Module):
    def __init__(self, in_nc=3, out_nc=3, n_feature=48, blindspot=False, zero_last=False):
        super().__init__()
        if blindspot:
            # arch_unet.py:196-198/:243-253 — out of the hot-path scope
            raise NotImplementedError("blindspot=True is not part of the B200 hot path")
        self.in_nc = in_nc
        self.out_nc = out_nc
        self.n_feature = n_feature
        self.blindspot = blindspot
        self.zero_last = zero_last
        nf = n_feature
        # same construction / initialisation order as arch_unet.py:114-192
        self.enc_conv0 = nn.Conv2d(in_nc, nf, 3, 1, 1)
        self.enc_conv1 = nn.Conv2d(nf, nf, 3, 1, 1)
        initialize_weights(self.enc_conv0, 0.1)
        initialize_weights(self.enc_conv1, 0.1)
        for i in range(2, 7):
            conv = nn.Conv2d(nf, nf, 3, 1, 1)
            setattr(self, f"enc_conv{i}", conv)
            initialize_weights(conv, 0.1)
        self.up5 = UpsampleCat(nf, nf)
        self.dec_conv5a = nn.Conv2d(nf * 2, nf * 2, 3, 1, 1)
        self.dec_conv5b = nn.Conv2d(nf * 2, nf * 2, 3, 1, 1)
        initialize_weights(self.dec_conv5a, 0.1)
        initialize_weights(self.dec_conv5b, 0.1)
        for lvl in (4, 3, 2):
            setattr(self, f"up{lvl}", UpsampleCat(nf * 2, nf * 2))
            a = nn.Conv2d(nf * 3, nf * 2, 3, 1, 1)
            b = nn.Conv2d(nf * 2, nf * 2, 3, 1, 1)
            setattr(self, f"dec_conv{lvl}a", a)
            setattr(self, f"dec_conv{lvl}b", b)
            initialize_weights(a, 0.1)
            initialize_weights(b, 0.1)
        self.up1 = UpsampleCat(nf * 2, nf * 2)
        self.dec_conv1a = nn.Conv2d(nf * 2 + in_nc, 96, 3, 1, 1)
        initialize_weights(self.dec_conv1a, 0.1)
        self.dec_conv1b = nn.Conv2d(96, 96, 3, 1, 1)
        initialize_weights(self.dec_conv1b, 0.1)
        self.nin_a = nn.Conv2d(96, 96, 1, 1, 0)
        self.nin_b = nn.Conv2d(96, 96, 1, 1, 0)
        initialize_weights(self.nin_a, 0.1)
        initialize_weights(self.nin_b, 0.1)
        self.nin_c = nn.Conv2d(96, out_nc, 1, 1, 0)
        if not self.zero_last:
            initialize_weights(self.nin_c, 0.1)
        self.precision = _ext.default_precision()
        self._cache = _PlanCache()
        self.last_launches = 0

    # -- engine glue ---------------------------------------------------------------------
    def set_precision(self, precision: str) -> "UNet":
        _ext.dtype_tag(precision)
        self.precision = precision
        return self

    def _plan_key(self, x, bwd: bool):
        n, c, h, w = x.shape
        return (self._arch, self.in_nc, self.out_nc, self.n_feature, n, h, w, _ext.dtype_tag(self.precision), bool(bwd))

    _arch = "unet"
    _num_params = 50
    _size_multiple = 32       # five 2x2 poolings (arch_unet.py:203-219)

    def _param_list(self):
        ps = list(self.parameters())
        if len(ps) != self._num_params:
            raise RuntimeError(f"expected {self._num_params} parameters, found {len(ps)}")
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("UNet parameters must be contiguous float32 tensors")
        return ps

    def forward(self, x):
        require_cuda(x, type(self).__name__ + ".forward")
        if x.dim() != 4 or x.shape[1] != self.in_nc:
            raise ValueError(f"expected input [N,{self.in_nc},H,W], got {tuple(x.shape)}")
        if x.shape[2] % self._size_multiple or x.shape[3] % self._size_multiple:
            raise ValueError("H and W must be multiples of 32 (five 2x2 poolings, arch_unet.py:203-219)")
        x = x.contiguous().float()
        params = self._param_list()
        for p in params:
            require_cuda(p, "UNet parameters")
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        if need_grad:
            return _UNetFunction.apply(self, x, *params)
        key = self._plan_key(x, False)
        plan, ws = self._cache.checkout(key, x.device)
        y = torch.empty((x.shape[0], self.out_nc, x.shape[2], x.shape[3]), dtype=torch.float32, device=x.device)
        check(lib().n2n_unet_forward(plan, ptr_array(params), ptr(x), ptr(y), ptr(ws), stream_ptr()))
        self.last_launches = lib().n2n_unet_launches(plan, 0)
        self._last = (plan, ws)
        self._cache.give_back(key, x.device, ws)
        return y

    ACTIVATION_BUFFERS = {"cat0": 0, "cat1": 1, "cat2": 2, "cat3": 3, "cat4": 4, "enc_conv0": 5, "enc_conv1": 6, "enc_conv2": 7,
                          "enc_conv3": 8, "enc_conv4": 9, "enc_conv5": 10, "pool5": 11, "enc_conv6": 12, "dec_conv5a": 13,
                          "dec_conv5b": 14, "dec_conv4a": 15, "dec_conv4b": 16, "dec_conv3a": 17, "dec_conv3b": 18,
                          "dec_conv2a": 19, "dec_conv2b": 20, "dec_conv1a": 21, "dec_conv1b": 22, "nin_a": 23, "nin_b": 24}

    def read_activation(self, name: str):
        """Layer-level parity hook: an intermediate tensor of the most recent no-grad forward as fp32 NCHW (all 16-channel
        blocks of the engine's buffer, zero / im2col padding included) — valid until the next forward of this module."""
        plan, ws = self._last
        dims = (_ext.ctypes.c_int * 3)()
        n = lib().n2n_unet_read_activation(plan, ptr(ws), self.ACTIVATION_BUFFERS[name], None, dims, stream_ptr())
        if n <= 0:
            raise ValueError(f"{name}: buffer not used by this plan")
        out = torch.empty((n // (dims[0] * dims[1] * dims[2]), dims[0], dims[1], dims[2]), dtype=torch.float32, device=ws.device)
        rc = lib().n2n_unet_read_activation(plan, ptr(ws), self.ACTIVATION_BUFFERS[name], ptr(out), dims, stream_ptr())
        if rc < 0:
            check(int(rc))
        return out


class RESNET(UNet):
    """Drop-in ``arch_unet.RESNET`` (reference arch_unet.py:263-409, non-blindspot): the UNet's 3x3 / 1x1 convolutions at
    full resolution (no pooling, no up-sampling), ``torch.cat([x, pool_k])`` skips and a global residual ``+ in_`` (:409).
    Same 42 parameters / names / registration and initialisation order as the reference — including ``up5.deconv.*``,
    which the reference constructs (:303) but never uses: it stays in the state_dict and receives no gradient.  Runs on
    the same native executor (n2n_resnet_plan_create + n2n_unet_forward / n2n_unet_backward)."""

    _arch = "resnet"
    _num_params = 42
    _size_multiple = 1
    _unused_params = (14, 15)          # up5.deconv.weight / .bias in parameter order

    def __init__(self, in_nc=3, out_nc=3, n_feature=48, blindspot=False, zero_last=False):
        nn.Module.__init__(self)
        if blindspot:
            raise NotImplementedError("blindspot=True is not part of the B200 hot path")
        if in_nc != out_nc:
            raise ValueError("RESNET adds its input to its output (arch_unet.py:409): in_nc must equal out_nc")
        self.in_nc = in_nc
        self.out_nc = out_nc
        self.n_feature = n_feature
        self.blindspot = blindspot
        self.zero_last = zero_last
        nf = n_feature
        # same construction / initialisation order as arch_unet.py:279-347
        self.enc_conv0 = nn.Conv2d(in_nc, nf, 3, 1, 1)
        self.enc_conv1 = nn.Conv2d(nf, nf, 3, 1, 1)
        initialize_weights(self.enc_conv0, 0.1)
        initialize_weights(self.enc_conv1, 0.1)
        for i in range(2, 7):
            conv = nn.Conv2d(nf, nf, 3, 1, 1)
            setattr(self, f"enc_conv{i}", conv)
            initialize_weights(conv, 0.1)
        self.up5 = UpsampleCat(nf, nf)
        self.dec_conv5a = nn.Conv2d(nf * 2, nf * 2, 3, 1, 1)
        self.dec_conv5b = nn.Conv2d(nf * 2, nf * 2, 3, 1, 1)
        initialize_weights(self.dec_conv5a, 0.1)
        initialize_weights(self.dec_conv5b, 0.1)
        for lvl in (4, 3, 2):
            a = nn.Conv2d(nf * 3, nf * 2, 3, 1, 1)
            b = nn.Conv2d(nf * 2, nf * 2, 3, 1, 1)
            setattr(self, f"dec_conv{lvl}a", a)
            setattr(self, f"dec_conv{lvl}b", b)
            initialize_weights(a, 0.1)
            initialize_weights(b, 0.1)
        self.dec_conv1a = nn.Conv2d(nf * 2 + in_nc, 96, 3, 1, 1)
        initialize_weights(self.dec_conv1a, 0.1)
        self.dec_conv1b = nn.Conv2d(96, 96, 3, 1, 1)
        initialize_weights(self.dec_conv1b, 0.1)
        self.nin_a = nn.Conv2d(96, 96, 1, 1, 0)
        self.nin_b = nn.Conv2d(96, 96, 1, 1, 0)
        initialize_weights(self.nin_a, 0.1)
        initialize_weights(self.nin_b, 0.1)
        self.nin_c = nn.Conv2d(96, out_nc, 1, 1, 0)
        if not self.zero_last:
            initialize_weights(self.nin_c, 0.1)
        self.precision = _ext.default_precision()
        self._cache = _PlanCache()
        self.last_launches = 0


from .improved import ImprovedUNet  # noqa: E402,F401  (arch_unet.py:475-531 — §8f N2, image_denoising_b200/improved.py)
