"""ctypes binding of libn2n_b200.so (the C-ABI declared in include/n2n_b200.h).

The library is built in-tree by ``build()`` (plain nvcc, sm_100a only) and loaded
lazily.  There is NO CPU fallback: if the shared object is missing, or a compute
entry point is called without a CUDA device, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libn2n_b200.so")
SOURCES = ["api.cu", "elementwise.cu", "subsample.cu", "loss_adam.cu", "metrics.cu", "pack.cu",
           "improved_ops.cu", "improved_plan.cu", "adapter_fused.cu", "tapgemm_simt.cu", "tapgemm_umma.cu", "slabgemm_umma.cu", "wgrad_slab_umma.cu", "head_umma.cu", "headbwd_umma.cu", "wgrad_umma.cu", "unet_plan.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

F32, BF16 = 0, 1
ADAM_CHUNK = 2048

_lib = None


def _newest_source_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(_HERE), "include")):
        for f in os.listdir(root):
            if os.path.isfile(os.path.join(root, f)):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libn2n_b200.so (in-tree).  Each translation unit
    is compiled to csrc/_obj/<name>.o in parallel (only when stale), then linked."""
    from concurrent.futures import ThreadPoolExecutor
    src_m = _newest_source_mtime()
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= src_m:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    hdr_m = max(os.path.getmtime(os.path.join(root, f))
                for root in (CSRC, os.path.join(os.path.dirname(_HERE), "include"))
                for f in os.listdir(root) if f.endswith((".h", ".cuh")))
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        spath = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(spath), hdr_m):
            return obj, None
        cmd = [nvcc] + flags + ["-c", "-o", obj, spath]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        return obj, (r.stdout + r.stderr if r.returncode != 0 else None)

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    errs = [e for _, e in results if e]
    if errs:
        raise RuntimeError("nvcc failed:\n" + "\n".join(errs))
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + [o for o, _ in results]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_SIGS = {
    "n2n_last_error": (c_char_p, []),
    "n2n_version": (c_int, []),
    "n2n_device_ok": (c_int, []),
    "n2n_launch_count": (ctypes.c_longlong, []),
    "n2n_profile_begin": (c_int, []),
    "n2n_profile_active": (c_int, []),
    "n2n_profile_end": (c_int, [POINTER(c_double)]),
    "n2n_profile_end_list": (c_int, [POINTER(c_double), c_int]),
    "n2n_mask_pair_from_rdidx": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "n2n_subsample": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "n2n_subsample_pair": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "n2n_space_to_depth": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "n2n_conv2d_workspace_bytes": (c_size_t, [c_int] * 7),
    "n2n_conv2d_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_float, c_int, c_void_p, c_void_p]),
    "n2n_conv2d_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_void_p, c_void_p]),
    "n2n_conv2d_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_void_p, c_void_p]),
    "n2n_deconv2x2_workspace_bytes": (c_size_t, [c_int] * 6),
    "n2n_deconv2x2_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_void_p, c_void_p]),
    "n2n_deconv2x2_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_void_p, c_void_p]),
    "n2n_deconv2x2_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_void_p, c_void_p]),
    "n2n_pool_workspace_bytes": (c_size_t, [c_int] * 5),
    "n2n_maxpool2_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "n2n_maxpool2_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int,
                                 c_void_p, c_void_p]),
    "n2n_unet_plan_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "n2n_resnet_plan_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "n2n_unet_plan_destroy": (None, [c_void_p]),
    "n2n_unet_workspace_bytes": (c_size_t, [c_void_p]),
    "n2n_unet_launches": (c_int, [c_void_p, c_int]),
    "n2n_unet_read_activation": (ctypes.c_longlong, [c_void_p, c_void_p, c_int, c_void_p, POINTER(c_int), c_void_p]),
    "n2n_unet_forward": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p]),
    "n2n_unet_share_weights": (c_int, [c_void_p, c_void_p, c_void_p]),
    "n2n_unet_pack_weights": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p]),
    "n2n_unet_backward": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, POINTER(c_void_p), c_void_p,
                                  c_void_p, c_void_p]),
    "n2n_adapter_plan_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "n2n_adapter_plan_destroy": (None, [c_void_p]),
    "n2n_adapter_workspace_bytes": (c_size_t, [c_void_p]),
    "n2n_adapter_forward": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "n2n_adapter_backward": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_void_p, c_void_p]),
    "n2n_loss_workspace_bytes": (c_size_t, [c_int64]),
    "n2n_loss_n2n_fwdbwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int64,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "n2n_loss_l1grad_fwdbwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "n2n_loss_structure_fwdbwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                          c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "n2n_loss_iqsl_workspace_bytes": (c_size_t, []),
    "n2n_loss_iqsl_fwdbwd": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_float, c_float, c_float,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "n2n_improved_plan_create": (c_int, [POINTER(c_void_p)] + [c_int] * 10),
    "n2n_improved_plan_destroy": (None, [c_void_p]),
    "n2n_improved_workspace_bytes": (c_size_t, [c_void_p]),
    "n2n_improved_num_params": (c_int, [c_void_p]),
    "n2n_improved_launches": (c_int, [c_void_p, c_int]),
    "n2n_improved_forward": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p]),
    "n2n_improved_read_buffer": (ctypes.c_longlong, [c_void_p, c_void_p, c_int, c_int, c_void_p, POINTER(c_int), c_void_p]),
    "n2n_improved_backward": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p, POINTER(c_void_p), c_void_p, c_void_p]),
    "n2n_groupnorm_groups": (c_int, [c_int, c_int]),
    "n2n_groupnorm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "n2n_groupnorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_float, c_float, c_void_p, c_void_p]),
    "n2n_groupnorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "n2n_act_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p]),
    "n2n_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p]),
    "n2n_add_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "n2n_pixel_shuffle2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "n2n_adam_multi": (c_int, [c_void_p, c_int, c_void_p, c_int, c_float, c_float, c_float, c_float, c_int,
                               c_float, c_void_p]),
    "n2n_set_step_scalars": (c_int, [c_void_p, c_float, c_float, c_float, c_float, c_int, c_void_p]),
    "n2n_loss_n2n_fwdbwd_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int64,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "n2n_adam_multi_dev": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_float, c_float, c_float,
                                   c_float, c_void_p]),
    "n2n_quantize_u8": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "n2n_tile_accumulate": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                    c_int, c_int, c_void_p]),
    "n2n_tile_finalize_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "n2n_tile_gather_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "n2n_tile_blend_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "n2n_psnr_ssim_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "n2n_psnr_ssim_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "n2n_crop_patches": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "n2n_probe_umma": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "n2n_debug_stall_buffer": (c_int, [c_void_p]),
    "n2n_probe_mma_rate": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
}

EXPORTS = tuple(_SIGS.keys())


def lib() -> ctypes.CDLL:
    """Load libn2n_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(image_denoising_b200 has no CPU / PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class N2NError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().n2n_last_error()
        raise N2NError(f"libn2n_b200 error {rc}: {msg.decode() if msg else '?'}")


def require_cuda(t, what: str):
    if not t.is_cuda:
        raise N2NError(f"{what}: tensor is on {t.device}; image_denoising_b200 runs on CUDA (sm_100a) only "
                       "and has no CPU fallback")


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def ptr_array(tensors):
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def dtype_tag(precision: str) -> int:
    p = precision.lower()
    if p in ("bf16", "bfloat16"):
        return BF16
    if p in ("fp32", "f32", "float32"):
        return F32
    raise ValueError(f"unknown precision {precision!r} (use 'bf16' or 'fp32')")


def default_precision() -> str:
    return os.environ.get("N2N_B200_PRECISION", "bf16")
