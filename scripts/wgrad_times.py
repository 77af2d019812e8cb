"""Weight-gradient engine timing on bench-shaped layers (diagnostic): full kernel vs MMAs skipped."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import _ext, ops
L = _ext.lib()
dev = torch.device("cuda:0")
N = int(os.environ.get("B", "64"))
cases = [("nin 96->96 1x1 @128^2", N, 96, 96, 128, 128, 1), ("d1b 96->96 3x3 @128^2", N, 96, 96, 128, 128, 3),
         ("enc1 48->48 3x3 @128^2", N, 48, 48, 128, 128, 3), ("d2a 144->96 3x3 @64^2", N, 144, 96, 64, 64, 3),
         ("nin_c 96->1 1x1 @128^2", N, 96, 1, 128, 128, 1), ("enc0-im2col 16->48 1x1 @128^2", N, 16, 48, 128, 128, 1),
         ("d1a-skip 16->96 1x1 @128^2", N, 16, 96, 128, 128, 1)]
for name, n, cin, cout, h, w, k in cases:
    x = torch.randn(n, cin, h, w, device=dev); dy = torch.randn(n, cout, h, w, device=dev)
    row = []
    for flags, ring in [(int(f), 0) for f in os.environ.get('FLAGS', '0,2,7').split(',')]:
        os.environ["N2N_DBG_FLAGS"] = str(flags); os.environ["N2N_WS_RING"] = str(ring) if ring else "9"
        for _ in range(2):
            ops.conv2d_wgrad(x, dy, k, "bf16")
        torch.cuda.synchronize()
        L.n2n_profile_begin()
        for _ in range(3):
            ops.conv2d_wgrad(x, dy, k, "bf16")
        out = (ctypes.c_double * 6)()
        L.n2n_profile_end(out)
        row.append(f"f{flags}/ring{ring or 'max'}: {out[3] / 3 * 1e3:7.1f}us")
    os.environ["N2N_DBG_FLAGS"] = "0"
    print(f"{name}: " + "  ".join(row))
