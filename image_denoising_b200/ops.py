"""Thin torch-tensor wrappers over the kernel-level C-ABI entry points (one Python
function per entry point of include/n2n_b200.h).  Everything here takes and returns
CUDA tensors; nothing falls back to PyTorch arithmetic."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _ext
from ._ext import check, lib, ptr, require_cuda, stream_ptr

_ELEM = {torch.float32: 4, torch.float16: 2, torch.bfloat16: 2, torch.float64: 8, torch.uint8: 1,
         torch.int8: 1, torch.int16: 2, torch.int32: 4, torch.int64: 8, torch.bool: 1}


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ----------------------------------------------------------------------------- sub-sampler
def mask_pair_from_rdidx(rd_idx: torch.Tensor, want_masks: bool = True, want_packed: bool = False):
    """train.py:151-172 from an already drawn rd_idx (int64, one value per cell)."""
    require_cuda(rd_idx, "mask_pair_from_rdidx")
    rd_idx = rd_idx.contiguous()
    cells = rd_idx.numel()
    m1 = m2 = pk = None
    if want_masks:
        m1 = torch.empty(cells * 4, dtype=torch.bool, device=rd_idx.device)
        m2 = torch.empty(cells * 4, dtype=torch.bool, device=rd_idx.device)
    if want_packed:
        pk = torch.empty(cells, dtype=torch.uint8, device=rd_idx.device)
    check(lib().n2n_mask_pair_from_rdidx(ptr(rd_idx), cells, ptr(m1), ptr(m2), ptr(pk), stream_ptr()))
    return m1, m2, pk


def subsample(img: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """train.py:175-190."""
    require_cuda(img, "generate_subimages")
    n, c, h, w = img.shape
    if mask.dtype != torch.bool or mask.dim() != 1 or mask.numel() != n * (h // 2) * (w // 2) * 4:
        raise ValueError(f"mask must be a 1-D bool tensor of length {n * (h // 2) * (w // 2) * 4}, got "
                         f"{tuple(mask.shape)} {mask.dtype}")
    img = img.contiguous(); mask = mask.contiguous()
    out = torch.empty((n, c, h // 2, w // 2), dtype=img.dtype, device=img.device)
    check(lib().n2n_subsample(ptr(img), ptr(mask), ptr(out), n, c, h, w, _ELEM[img.dtype], stream_ptr()))
    return out


def subsample_pair(img: torch.Tensor, mask1: Optional[torch.Tensor] = None, mask2: Optional[torch.Tensor] = None,
                   packed: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused form: both sub-images in one pass over img."""
    require_cuda(img, "subsample_pair")
    n, c, h, w = img.shape
    cells = n * (h // 2) * (w // 2)
    if packed is None:
        for m in (mask1, mask2):
            if m is None or m.dtype != torch.bool or m.numel() != cells * 4:
                raise ValueError("subsample_pair needs two bool masks of 4*cells elements or a packed selector")
        mask1 = mask1.contiguous(); mask2 = mask2.contiguous()
    elif packed.dtype != torch.uint8 or packed.numel() != cells:
        raise ValueError("packed selector must be uint8 with one byte per cell")
    img = img.contiguous()
    o1 = torch.empty((n, c, h // 2, w // 2), dtype=img.dtype, device=img.device)
    o2 = torch.empty_like(o1)
    check(lib().n2n_subsample_pair(ptr(img), ptr(mask1), ptr(mask2), ptr(packed), ptr(o1), ptr(o2),
                                   n, c, h, w, _ELEM[img.dtype], stream_ptr()))
    return o1, o2


def space_to_depth(x: torch.Tensor, block_size: int) -> torch.Tensor:
    """train.py:134-138: [n,c,h,w] -> [n, c*bs*bs, h//bs, w//bs], channel = c*bs*bs + ky*bs + kx."""
    require_cuda(x, "space_to_depth")
    n, c, h, w = x.shape
    bs = int(block_size)
    if bs < 1 or h % bs or w % bs:
        raise ValueError(f"space_to_depth: H, W = {h}, {w} must be multiples of block_size {block_size}")
    x = x.contiguous()
    y = torch.empty((n, c * bs * bs, h // bs, w // bs), dtype=x.dtype, device=x.device)
    check(lib().n2n_space_to_depth(ptr(x), ptr(y), n, c, h, w, bs, _ELEM[x.dtype], stream_ptr()))
    return y


# ----------------------------------------------------------------------------- single layers
def conv2d_fwd(x, w, b=None, act_slope: float = -1.0, precision: str = "fp32"):
    require_cuda(x, "conv2d_fwd")
    x = _f32c(x); w = _f32c(w); b = None if b is None else _f32c(b)
    n, cin, h, wd = x.shape
    cout, _, k, _ = w.shape
    dt = _ext.dtype_tag(precision)
    y = torch.empty((n, cout, h, wd), dtype=torch.float32, device=x.device)
    ws = _ws(lib().n2n_conv2d_workspace_bytes(n, cin, cout, h, wd, k, dt), x.device)
    check(lib().n2n_conv2d_fwd(ptr(x), ptr(w), ptr(b), ptr(y), n, cin, cout, h, wd, k, act_slope, dt, ptr(ws), stream_ptr()))
    return y


def conv2d_dgrad(dy, w, precision: str = "fp32"):
    require_cuda(dy, "conv2d_dgrad")
    dy = _f32c(dy); w = _f32c(w)
    n, cout, h, wd = dy.shape
    _, cin, k, _ = w.shape
    dt = _ext.dtype_tag(precision)
    dx = torch.empty((n, cin, h, wd), dtype=torch.float32, device=dy.device)
    ws = _ws(lib().n2n_conv2d_workspace_bytes(n, cin, cout, h, wd, k, dt), dy.device)
    check(lib().n2n_conv2d_dgrad(ptr(dy), ptr(w), ptr(dx), n, cin, cout, h, wd, k, dt, ptr(ws), stream_ptr()))
    return dx


def conv2d_wgrad(x, dy, ksize: int, precision: str = "fp32"):
    require_cuda(x, "conv2d_wgrad")
    x = _f32c(x); dy = _f32c(dy)
    n, cin, h, wd = x.shape
    cout = dy.shape[1]
    dt = _ext.dtype_tag(precision)
    dw = torch.empty((cout, cin, ksize, ksize), dtype=torch.float32, device=x.device)
    db = torch.empty((cout,), dtype=torch.float32, device=x.device)
    ws = _ws(lib().n2n_conv2d_workspace_bytes(n, cin, cout, h, wd, ksize, dt), x.device)
    check(lib().n2n_conv2d_wgrad(ptr(x), ptr(dy), ptr(dw), ptr(db), n, cin, cout, h, wd, ksize, dt, ptr(ws), stream_ptr()))
    return dw, db


def deconv2x2_fwd(x, w, b=None, precision: str = "fp32"):
    require_cuda(x, "deconv2x2_fwd")
    x = _f32c(x); w = _f32c(w); b = None if b is None else _f32c(b)
    n, cin, h, wd = x.shape
    cout = w.shape[1]
    dt = _ext.dtype_tag(precision)
    y = torch.empty((n, cout, 2 * h, 2 * wd), dtype=torch.float32, device=x.device)
    ws = _ws(lib().n2n_deconv2x2_workspace_bytes(n, cin, cout, h, wd, dt), x.device)
    check(lib().n2n_deconv2x2_fwd(ptr(x), ptr(w), ptr(b), ptr(y), n, cin, cout, h, wd, dt, ptr(ws), stream_ptr()))
    return y


def deconv2x2_dgrad(dy, w, precision: str = "fp32"):
    require_cuda(dy, "deconv2x2_dgrad")
    dy = _f32c(dy); w = _f32c(w)
    n, cout, h2, w2 = dy.shape
    cin = w.shape[0]
    h, wd = h2 // 2, w2 // 2
    dt = _ext.dtype_tag(precision)
    dx = torch.empty((n, cin, h, wd), dtype=torch.float32, device=dy.device)
    ws = _ws(lib().n2n_deconv2x2_workspace_bytes(n, cin, cout, h, wd, dt), dy.device)
    check(lib().n2n_deconv2x2_dgrad(ptr(dy), ptr(w), ptr(dx), n, cin, cout, h, wd, dt, ptr(ws), stream_ptr()))
    return dx


def deconv2x2_wgrad(x, dy, precision: str = "fp32"):
    require_cuda(x, "deconv2x2_wgrad")
    x = _f32c(x); dy = _f32c(dy)
    n, cin, h, wd = x.shape
    cout = dy.shape[1]
    dt = _ext.dtype_tag(precision)
    dw = torch.empty((cin, cout, 2, 2), dtype=torch.float32, device=x.device)
    db = torch.empty((cout,), dtype=torch.float32, device=x.device)
    ws = _ws(lib().n2n_deconv2x2_workspace_bytes(n, cin, cout, h, wd, dt), x.device)
    check(lib().n2n_deconv2x2_wgrad(ptr(x), ptr(dy), ptr(dw), ptr(db), n, cin, cout, h, wd, dt, ptr(ws), stream_ptr()))
    return dw, db


def maxpool2_fwd(x, precision: str = "fp32"):
    require_cuda(x, "maxpool2_fwd")
    x = _f32c(x)
    n, c, h, w = x.shape
    dt = _ext.dtype_tag(precision)
    y = torch.empty((n, c, h // 2, w // 2), dtype=torch.float32, device=x.device)
    ws = _ws(lib().n2n_pool_workspace_bytes(n, c, h, w, dt), x.device)
    check(lib().n2n_maxpool2_fwd(ptr(x), ptr(y), n, c, h, w, dt, ptr(ws), stream_ptr()))
    return y


def maxpool2_bwd(x, dy, slope: float = 1.0, precision: str = "fp32"):
    require_cuda(x, "maxpool2_bwd")
    x = _f32c(x); dy = _f32c(dy)
    n, c, h, w = x.shape
    dt = _ext.dtype_tag(precision)
    dx = torch.empty_like(x)
    ws = _ws(lib().n2n_pool_workspace_bytes(n, c, h, w, dt), x.device)
    check(lib().n2n_maxpool2_bwd(ptr(x), ptr(dy), ptr(dx), n, c, h, w, slope, dt, ptr(ws), stream_ptr()))
    return dx


# ----------------------------------------------------------------------------- losses
_loss_ws = {}


def _loss_workspace(device) -> torch.Tensor:
    key = (device.type, device.index)
    if key not in _loss_ws:
        _loss_ws[key] = torch.zeros(lib().n2n_loss_workspace_bytes(0), dtype=torch.uint8, device=device)
    return _loss_ws[key]


def n2n_loss_fwdbwd(out, sub2, den1, den2, lam: float, grad_scale: float = 1.0, want_grad: bool = True):
    """training_script.md:146-153 -> (loss3 tensor [loss_all, loss1, loss2], dloss/dout or None)."""
    require_cuda(out, "n2n_loss")
    out = _f32c(out); sub2 = _f32c(sub2); den1 = _f32c(den1); den2 = _f32c(den2)
    if not (out.shape == sub2.shape == den1.shape == den2.shape):
        raise ValueError("n2n_loss: shape mismatch")
    loss3 = torch.empty(3, dtype=torch.float32, device=out.device)
    grad = torch.empty_like(out) if want_grad else None
    check(lib().n2n_loss_n2n_fwdbwd(ptr(out), ptr(sub2), ptr(den1), ptr(den2), float(lam), float(grad_scale),
                                    out.numel(), ptr(loss3), ptr(grad), ptr(_loss_workspace(out.device)), stream_ptr()))
    return loss3, grad


def l1grad_loss_fwdbwd(pred, target, lambda_grad: float, grad_scale: float = 1.0, want_grad: bool = True):
    """finetune.py:153-162, :283-285 -> (loss3 [loss, l1, grad_term], dloss/dpred or None)."""
    require_cuda(pred, "l1grad_loss")
    pred = _f32c(pred); target = _f32c(target)
    if pred.shape != target.shape or pred.dim() != 4:
        raise ValueError("l1grad_loss: need two [N,C,H,W] tensors of equal shape")
    n, c, h, w = pred.shape
    loss3 = torch.empty(3, dtype=torch.float32, device=pred.device)
    grad = torch.empty_like(pred) if want_grad else None
    check(lib().n2n_loss_l1grad_fwdbwd(ptr(pred), ptr(target), n, c, h, w, float(lambda_grad), float(grad_scale),
                                       ptr(loss3), ptr(grad), ptr(_loss_workspace(pred.device)), stream_ptr()))
    return loss3, grad


def structure_loss_fwdbwd(pred, pred2, target, alpha: float, beta: float, gamma: float, grad_scale: float = 1.0,
                          want_grad: bool = True):
    """util.py:41-70 -> (loss4 [loss, pixel, TV, consistency], dloss/dpred or None, dloss/dpred2 or None)."""
    require_cuda(pred, "structure_loss")
    pred = _f32c(pred); pred2 = _f32c(pred2); target = _f32c(target)
    if not (pred.shape == pred2.shape == target.shape) or pred.dim() != 4:
        raise ValueError("structure_loss: need three [N,C,H,W] tensors of equal shape")
    n, c, h, w = pred.shape
    loss4 = torch.empty(4, dtype=torch.float32, device=pred.device)
    g1 = torch.empty_like(pred) if want_grad else None
    g2 = torch.empty_like(pred2) if want_grad else None
    check(lib().n2n_loss_structure_fwdbwd(ptr(pred), ptr(pred2), ptr(target), n, c, h, w, float(alpha), float(beta),
                                          float(gamma), float(grad_scale), ptr(loss4), ptr(g1), ptr(g2),
                                          ptr(_loss_workspace(pred.device)), stream_ptr()))
    return loss4, g1, g2


_iqsl_ws = {}


def iqsl_loss_fwdbwd(pred, target, t1: float, t2: float, tau: float = 0.1, margin: float = 0.0, ce_factor: float = 0.5,
                     eps: float = 1e-6, grad_scale: float = 1.0, want_grad: bool = True):
    """finetune_iqsl.py:291-383 -> (loss3 [total, dice, ce], dloss/dpred or None); single-channel tensors in [0,1]."""
    require_cuda(pred, "iqsl_loss")
    pred = _f32c(pred); target = _f32c(target)
    if pred.shape != target.shape:
        raise ValueError("pred and target must have the same shape.")
    if pred.dim() == 4 and pred.shape[1] != 1:
        raise ValueError("IQSL currently assumes single-channel grayscale input.")
    key = (pred.device.type, pred.device.index)
    if key not in _iqsl_ws:
        _iqsl_ws[key] = torch.zeros(lib().n2n_loss_iqsl_workspace_bytes(), dtype=torch.uint8, device=pred.device)
    loss3 = torch.empty(3, dtype=torch.float32, device=pred.device)
    grad = torch.empty_like(pred) if want_grad else None
    check(lib().n2n_loss_iqsl_fwdbwd(ptr(pred), ptr(target), pred.numel(), float(t1), float(t2), float(tau), float(margin),
                                     float(ce_factor), float(eps), float(grad_scale), ptr(loss3), ptr(grad), ptr(_iqsl_ws[key]),
                                     stream_ptr()))
    return loss3, grad


# ----------------------------------------------------------------------------- ImprovedUNet operators (arch_unet.py:420-531)
def groupnorm_groups(channels: int, groups: int = 32) -> int:
    """arch_unet.py:11-15: the largest g <= min(groups, channels) that divides channels."""
    return int(lib().n2n_groupnorm_groups(int(channels), int(groups)))


def groupnorm_fwd(x, gamma, beta, groups: int, eps: float = 1e-5, act_slope: float = -1.0, residual=None, want_stats: bool = True):
    """-> (y, mean_rstd [n, groups, 2] or None).  y = GN(x)*gamma + beta, then LeakyReLU(act_slope) if act_slope >= 0, + residual if given."""
    require_cuda(x, "groupnorm_fwd")
    x = _f32c(x); gamma = _f32c(gamma); beta = _f32c(beta)
    residual = None if residual is None else _f32c(residual)
    n, c, h, w = x.shape
    y = torch.empty_like(x)
    stats = torch.empty((n, groups, 2), dtype=torch.float32, device=x.device) if want_stats else None
    ws = _ws(lib().n2n_groupnorm_workspace_bytes(n, c), x.device)
    check(lib().n2n_groupnorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(residual), ptr(y), ptr(stats), n, c, h * w, int(groups),
                                  float(eps), float(act_slope), ptr(ws), stream_ptr()))
    return y, stats


def groupnorm_bwd(x, gamma, y, dy, stats, groups: int, act_slope: float = -1.0):
    """-> (dx, dgamma, dbeta); ``y`` (forward output) is only read when an activation was fused."""
    require_cuda(x, "groupnorm_bwd")
    x = _f32c(x); gamma = _f32c(gamma); dy = _f32c(dy)
    n, c, h, w = x.shape
    dx = torch.empty_like(x)
    dgamma = torch.empty_like(gamma); dbeta = torch.empty_like(gamma)
    ws = _ws(lib().n2n_groupnorm_workspace_bytes(n, c), x.device)
    check(lib().n2n_groupnorm_bwd(ptr(x), ptr(gamma), ptr(y), ptr(dy), ptr(stats), ptr(dx), ptr(dgamma), ptr(dbeta), n, c, h * w,
                                  int(groups), float(act_slope), ptr(ws), stream_ptr()))
    return dx, dgamma, dbeta


ACT_LRELU, ACT_SIGMOID = 1, 2


def act_fwd(x, kind: int, slope: float = 0.2):
    require_cuda(x, "act_fwd")
    x = _f32c(x)
    y = torch.empty_like(x)
    check(lib().n2n_act_fwd(ptr(x), ptr(y), x.numel(), int(kind), float(slope), stream_ptr()))
    return y


def act_bwd(y, dy, kind: int, slope: float = 0.2):
    require_cuda(y, "act_bwd")
    y = _f32c(y); dy = _f32c(dy)
    dx = torch.empty_like(y)
    check(lib().n2n_act_bwd(ptr(y), ptr(dy), ptr(dx), y.numel(), int(kind), float(slope), stream_ptr()))
    return dx


def add(a, b):
    require_cuda(a, "add")
    a = _f32c(a); b = _f32c(b)
    if a.shape != b.shape:
        raise ValueError("add: shape mismatch")
    out = torch.empty_like(a)
    check(lib().n2n_add_f32(ptr(a), ptr(b), ptr(out), a.numel(), stream_ptr()))
    return out


def pixel_shuffle2(x, inverse: bool = False):
    """nn.PixelShuffle(2) ([n,4c,h,w] -> [n,c,2h,2w]) or, with inverse=True, its gradient map ([n,c,2h,2w] -> [n,4c,h,w])."""
    require_cuda(x, "pixel_shuffle2")
    x = _f32c(x)
    n, c, h, w = x.shape
    if inverse:
        if h % 2 or w % 2:
            raise ValueError("pixel_shuffle2(inverse): H and W must be even")
        y = torch.empty((n, 4 * c, h // 2, w // 2), dtype=torch.float32, device=x.device)
        check(lib().n2n_pixel_shuffle2(ptr(x), ptr(y), n, c, h // 2, w // 2, 1, stream_ptr()))
    else:
        if c % 4:
            raise ValueError("pixel_shuffle2: channels must be a multiple of 4")
        y = torch.empty((n, c // 4, 2 * h, 2 * w), dtype=torch.float32, device=x.device)
        check(lib().n2n_pixel_shuffle2(ptr(x), ptr(y), n, c // 4, h, w, 0, stream_ptr()))
    return y


# ----------------------------------------------------------------------------- evaluation
def quantize_u8(pred: torch.Tensor, bias: float) -> torch.Tensor:
    require_cuda(pred, "quantize_u8")
    pred = _f32c(pred)
    out = torch.empty(pred.shape, dtype=torch.uint8, device=pred.device)
    check(lib().n2n_quantize_u8(ptr(pred), ptr(out), pred.numel(), float(bias), stream_ptr()))
    return out


def tile_accumulate(pred_tile, weight_mask, acc, cnt, r0: int, c0: int, th: int, tw: int):
    ps = pred_tile.shape[-1]
    H, W = acc.shape
    check(lib().n2n_tile_accumulate(ptr(pred_tile), ps, ptr(weight_mask), ptr(acc), ptr(cnt), H, W, r0, c0, th, tw,
                                    stream_ptr()))


def tile_finalize_u8(acc, cnt) -> torch.Tensor:
    out = torch.empty(acc.shape, dtype=torch.uint8, device=acc.device)
    check(lib().n2n_tile_finalize_u8(ptr(acc), ptr(cnt), ptr(out), acc.numel(), stream_ptr()))
    return out


def tile_gather_u8(images: torch.Tensor, ps: int, stride: int) -> torch.Tensor:
    """uint8 [B,H,W] (CUDA) -> fp32 tiles [B*T,1,ps,ps] (evaluation_704.py:82-96: cut, /255, reflect-pad)."""
    require_cuda(images, "tile_gather_u8")
    if images.dtype != torch.uint8 or images.dim() != 3:
        raise ValueError("tile_gather_u8 takes a uint8 [B,H,W] tensor")
    images = images.contiguous()
    b, h, w = images.shape
    t = ((h + stride - 1) // stride) * ((w + stride - 1) // stride)
    tiles = torch.empty((b * t, 1, ps, ps), dtype=torch.float32, device=images.device)
    check(lib().n2n_tile_gather_u8(ptr(images), b, h, w, ps, stride, ptr(tiles), stream_ptr()))
    return tiles


def tile_blend_u8(pred_tiles: torch.Tensor, weight_mask: torch.Tensor, batch: int, h: int, w: int, ps: int, stride: int) -> torch.Tensor:
    """fp32 tiles [B*T,1,ps,ps] -> uint8 [B,H,W] (evaluation_704.py:103-120: triangular blend, truncating quantiser)."""
    require_cuda(pred_tiles, "tile_blend_u8")
    pred_tiles = _f32c(pred_tiles); weight_mask = _f32c(weight_mask)
    out = torch.empty((batch, h, w), dtype=torch.uint8, device=pred_tiles.device)
    check(lib().n2n_tile_blend_u8(ptr(pred_tiles), ptr(weight_mask), batch, h, w, ps, stride, ptr(out), stream_ptr()))
    return out


def psnr_ssim_u8(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a, b: uint8 CUDA tensors [B,H,W] or [B,H,W,C] (C in {1,3}) -> float64 [B,2] = (psnr, ssim)."""
    require_cuda(a, "psnr_ssim")
    if a.shape != b.shape or a.dtype != torch.uint8 or b.dtype != torch.uint8:
        raise ValueError("Input images must have the same dimensions.")
    if a.dim() == 3:
        a = a.unsqueeze(-1); b = b.unsqueeze(-1)
    a = a.contiguous(); b = b.contiguous()
    B, H, W, C = a.shape
    res = torch.empty((B, 2), dtype=torch.float64, device=a.device)
    ws = _ws(lib().n2n_psnr_ssim_workspace_bytes(B, H, W, C), a.device)
    check(lib().n2n_psnr_ssim_u8(ptr(a), ptr(b), B, H, W, C, ptr(res), ptr(ws), stream_ptr()))
    return res
