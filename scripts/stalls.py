"""Where does the tap-GEMM pipeline wait?  Runs single conv layers (bench shapes) with the stall counters on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_denoising_b200 import _ext, ops
L = _ext.lib()
dev = torch.device("cuda:0")
dbg = torch.zeros(8, dtype=torch.int64, device=dev)
cases = [("d1b 96->96 3x3 @256^2 x16", 16, 96, 96, 256, 256, 3), ("d1a 97->96", 16, 97, 96, 256, 256, 3),
         ("nin 96->96 1x1", 16, 96, 96, 256, 256, 1), ("enc1 48->48", 16, 48, 48, 256, 256, 3)]
for name, n, cin, cout, h, w, k in cases:
    x = torch.randn(n, cin, h, w, device=dev); wt = torch.randn(cout, cin, k, k, device=dev) * 0.05
    b = torch.zeros(cout, device=dev)
    for _ in range(2):
        ops.conv2d_fwd(x, wt, b, 0.2, "bf16")
    L.n2n_debug_stall_buffer(dbg.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv2d_fwd(x, wt, b, 0.2, "bf16"); e1.record(); torch.cuda.synchronize()
    L.n2n_debug_stall_buffer(None)
    d = dbg.cpu().tolist()
    tiles = max(d[5], 1)
    print(f"{name}: whole op {e0.elapsed_time(e1)*1e3:.0f} us (incl. layout conversions); CTA0 tiles={tiles}")
    print(f"   per tile cycles: producer wait-empty {d[0]/tiles:.0f} / total {d[1]/tiles:.0f}; mma wait-data {d[2]/tiles:.0f} "
          f"wait-accum {d[3]/tiles:.0f} / total {d[4]/tiles:.0f}; epilogue wait {d[6]/tiles:.0f} / total {d[7]/tiles:.0f}")
