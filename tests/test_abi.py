"""CPU: the C-ABI library builds, loads, and exports exactly the symbols include/n2n_b200.h
declares (no compute calls without a GPU); host-side logic that needs no device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "n2n_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(n2n_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from image_denoising_b200 import _ext
    _ext.build()
    L = _ext.lib()
    names = _header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/n2n_b200.h but not exported"
    assert sorted(_ext.EXPORTS) == names, "ctypes signature table and header disagree"
    assert L.n2n_version() >= 100
    assert L.n2n_last_error() is not None


def test_compute_fails_loudly_without_cuda():
    from image_denoising_b200 import UNet, _ext, generate_subimages
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    assert _ext.lib().n2n_device_ok() == 0
    with pytest.raises(_ext.N2NError):
        UNet(1, 1, 4)(torch.zeros(1, 1, 32, 32))
    with pytest.raises(_ext.N2NError):
        generate_subimages(torch.zeros(1, 1, 4, 4), torch.zeros(16, dtype=torch.bool))


def test_argument_validation_needs_no_device():
    from image_denoising_b200 import _ext
    L = _ext.lib()
    h = ctypes.c_void_p()
    assert L.n2n_unet_plan_create(ctypes.byref(h), 1, 1, 48, 1, 100, 64, 0, 0) == -1     # H not a multiple of 32
    assert b"multiples of 32" in L.n2n_last_error()
    assert L.n2n_unet_plan_create(ctypes.byref(h), 1, 1, 48, 4, 256, 256, 1, 1) == 0
    assert L.n2n_unet_workspace_bytes(h) > 0
    L.n2n_unet_plan_destroy(h)
    assert L.n2n_subsample(None, None, None, 1, 1, 4, 4, 3, None) == -1                  # NULL pointers


def test_unet_state_dict_layout_and_init_stream():
    """Same keys / shapes / order as the oracle's inventory (arch_unet.py:114-192); 1 256 689 params."""
    from image_denoising_b200 import UNet
    from oracle import n2n_oracle as O
    net = UNet(in_nc=1, out_nc=1, n_feature=48)
    shapes = O.unet_param_shapes(1, 1, 48)
    sd = net.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(v.shape) == shapes[k] for k, v in sd.items())
    assert sum(v.numel() for v in sd.values()) == 1256689
    assert all(float(v.abs().max()) == 0.0 for k, v in sd.items() if k.endswith(".bias"))
    ref_dir = os.environ.get("N2N_REFERENCE", "/root/reference")
    if os.path.exists(os.path.join(ref_dir, "arch_unet.py")):
        # in the build container: identical RNG stream as the reference constructor
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_arch_unet", os.path.join(ref_dir, "arch_unet.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        torch.manual_seed(1234)
        a = ref.UNet(in_nc=3, out_nc=3, n_feature=8).state_dict()
        torch.manual_seed(1234)
        b = UNet(in_nc=3, out_nc=3, n_feature=8).state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_adapter_state_dict_layout():
    from image_denoising_b200 import DenoiserWithAdapter, UNet
    m = DenoiserWithAdapter(UNet(3, 3, 8), in_channels=3, hidden_channels=16)
    keys = list(m.state_dict().keys())
    assert len(keys) == 54 and keys[-4:] == ["adapter.net.0.weight", "adapter.net.0.bias",
                                             "adapter.net.2.weight", "adapter.net.2.bias"]
    assert tuple(m.state_dict()["adapter.net.0.weight"].shape) == (16, 6, 3, 3)
    assert all(not p.requires_grad for p in m.base.parameters())


def test_tile_weight_and_origins_match_oracle():
    from image_denoising_b200 import evaluate
    from oracle import n2n_oracle as O
    assert np.array_equal(evaluate.tile_weight(352), O.tile_weight(352))
    assert evaluate.tile_origins(704, 704) == [(r, c) for r in (0, 288, 576) for c in (0, 288, 576)]


def test_resnet_state_dict_layout_and_init_stream():
    """arch_unet.RESNET (arch_unet.py:263-347): 42 tensors in the reference's order, up5.deconv.* included."""
    from image_denoising_b200 import RESNET
    from oracle import n2n_oracle as O
    net = RESNET(in_nc=1, out_nc=1, n_feature=8)
    shapes = O.resnet_param_shapes(1, 1, 8)
    sd = net.state_dict()
    assert list(sd.keys()) == list(shapes.keys()) and len(sd) == 42
    assert all(tuple(v.shape) == shapes[k] for k, v in sd.items())
    with pytest.raises(ValueError):
        RESNET(in_nc=1, out_nc=3)
    ref_dir = os.environ.get("N2N_REFERENCE", "/root/reference")
    if os.path.exists(os.path.join(ref_dir, "arch_unet.py")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_arch_unet2", os.path.join(ref_dir, "arch_unet.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        torch.manual_seed(7)
        a = ref.RESNET(in_nc=3, out_nc=3, n_feature=8).state_dict()
        torch.manual_seed(7)
        b = RESNET(in_nc=3, out_nc=3, n_feature=8).state_dict()
        assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)
