"""Bring-up probe for the tcgen05 descriptor encodings (see n2n_probe_umma in include/n2n_b200.h):
a bare single-CTA GEMM whose operands are laid out in shared memory exactly as the conv engines
assume.  Base 0 = K-major SWIZZLE_32B (forward / dgrad engine), base 1 = MN-major SWIZZLE_32B
(weight-gradient engine); bases 3 / 4 start the operand `shift` 32-byte rows into a larger staged
slab (tap-shifted views of one slab)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(variant, a, b, n, k):
    from image_denoising_b200._ext import check, lib, ptr, stream_ptr
    dev = torch.device("cuda:0")
    d = torch.full((128, n), float("nan"), device=dev)
    ad, bd = a.to(dev).contiguous(), b.to(dev).contiguous()
    check(lib().n2n_probe_umma(variant, ptr(ad), ptr(bd), ptr(d), 128, n, k, stream_ptr()))
    torch.cuda.synchronize()
    return d.cpu()


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,k", [(96, 48), (48, 16), (144, 128), (16, 32), (256, 64)])
def test_probe_umma(variant, n, k):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g = torch.Generator().manual_seed(n * 1000 + k)
    a = torch.randn(128, k, generator=g).bfloat16()
    b = torch.randn(n, k, generator=g).bfloat16()
    ref = a.float() @ b.float().t()
    d = _run(variant, a, b, n, k)
    err = (d - ref).abs().max().item()
    assert err < 1e-2 * max(1.0, ref.abs().max().item()), f"variant {variant} n={n} k={k}: max err {err}"


def test_probe_shifted_starts_report():
    """Diagnostic (never fails): which descriptor recipe makes a row-shifted start address work."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n, k = 96, 48
    g = torch.Generator().manual_seed(7)
    lines = []
    for base in (3, 4):
        for use_bo in (0, 1):
            for shift in range(0, 9):
                if base == 3:
                    a = torch.randn(136, k, generator=g).bfloat16()
                    b = torch.randn(n, k, generator=g).bfloat16()
                    ref = a[shift:shift + 128].float() @ b.float().t()
                else:
                    a = torch.randn(128, k, generator=g).bfloat16()          # [m][k]
                    b = torch.randn(n, k + 8, generator=g).bfloat16()        # [n][k + 8]
                    ref = a.float() @ b[:, shift:shift + k].float().t()
                d = _run(base | (shift << 8) | (use_bo << 16), a, b, n, k)
                err = (d - ref).abs().max().item() / max(1.0, ref.abs().max().item())
                lines.append(f"PROBE base={base} base_offset={use_bo} shift={shift} rel_err={err:.3e} {'OK' if err < 1e-2 else 'BAD'}")
    print("\n" + "\n".join(lines))
