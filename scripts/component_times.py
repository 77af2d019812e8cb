"""Which pipeline role bounds the tap-GEMM?  Times single conv layers (bench shapes) with parts of the
kernel switched off (N2N_DBG_FLAGS: 1 = no TMA loads, 2 = no MMA issue, 4 = no epilogue stores)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import _ext, ops
L = _ext.lib()
dev = torch.device("cuda:0")
N = int(os.environ.get("B", "32"))
cases = [("d1b 96->96 3x3 @256^2", N, 96, 96, 256, 256, 3), ("d1a 97->96 3x3 @256^2", N, 97, 96, 256, 256, 3),
         ("nin 96->96 1x1 @256^2", N, 96, 96, 256, 256, 1), ("enc1 48->48 3x3 @256^2", N, 48, 48, 256, 256, 3),
         ("d2a 144->96 3x3 @128^2", N, 144, 96, 128, 128, 3),
         ("d5a 96->96 3x3 @16^2 x64", 64, 96, 96, 16, 16, 3), ("enc5 48->48 3x3 @16^2 x64", 64, 48, 48, 16, 16, 3),
         ("d4b 96->96 3x3 @32^2 x64", 64, 96, 96, 32, 32, 3)]
if os.environ.get("CASES"):
    cases = [cases[int(i)] for i in os.environ["CASES"].split(",")]
for name, n, cin, cout, h, w, k in cases:
    x = torch.randn(n, cin, h, w, device=dev); wt = torch.randn(cout, cin, k, k, device=dev) * 0.05
    b = torch.zeros(cout, device=dev)
    flops = 2.0 * n * h * w * cout * cin * k * k
    row = []
    for flags in [int(f) for f in os.environ.get("FLAGS", "0,1,2,4,3,5,6,7").split(",")]:
        os.environ["N2N_DBG_FLAGS"] = str(flags)
        for _ in range(2):
            ops.conv2d_fwd(x, wt, b, 0.2, "bf16")
        torch.cuda.synchronize()
        L.n2n_profile_begin()
        for _ in range(3):
            ops.conv2d_fwd(x, wt, b, 0.2, "bf16")
        out = (ctypes.c_double * 6)()
        L.n2n_profile_end(out)
        ms = out[0] / 3
        row.append(f"f{flags}:{ms*1e3:7.1f}us")
    os.environ["N2N_DBG_FLAGS"] = "0"
    print(f"{name}: " + " ".join(row) + f"   (full = {flops/1e9:.0f} GFLOP -> {flops/(float(row[0].split(':')[1][:-2])*1e-6)/1e12:.0f} TF/s)")
