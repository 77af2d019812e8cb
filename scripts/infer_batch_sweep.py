import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
class D: pass
for pl in (8, 16, 32):
    r = bench.inference_704(dev, "bf16", 1, 0, None, total_images=128, per_launch=pl, reps=2)
    t = bench.inference_704_tiled(dev, "bf16", 1, 0, None, total_images=128, per_launch=pl, reps=1)
    print(pl, round(r["value"], 1), round(t["value"], 1))
    torch.cuda.empty_cache()
