"""Bring-up probe for the tcgen05 descriptor encodings (see n2n_probe_umma in include/n2n_b200.h):
a bare single-CTA GEMM whose operands are laid out in shared memory exactly as the conv engines
assume.  Variant 0 = K-major SWIZZLE_32B (forward / dgrad engine), variant 1 = MN-major
SWIZZLE_32B (weight-gradient engine)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,k", [(96, 48), (48, 16), (144, 128), (16, 32), (256, 64)])
def test_probe_umma(variant, n, k):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from image_denoising_b200._ext import check, lib, ptr, stream_ptr
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n * 1000 + k)
    a = torch.randn(128, k, generator=g).bfloat16()
    b = torch.randn(n, k, generator=g).bfloat16()
    ref = a.float() @ b.float().t()
    d = torch.full((128, n), float("nan"), device=dev)
    ad, bd = a.to(dev), b.to(dev)
    check(lib().n2n_probe_umma(variant, ptr(ad), ptr(bd), ptr(d), 128, n, k, stream_ptr()))
    torch.cuda.synchronize()
    err = (d.cpu() - ref).abs().max().item()
    assert err < 1e-2 * max(1.0, ref.abs().max().item()), f"variant {variant} n={n} k={k}: max err {err}"
