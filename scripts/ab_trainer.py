"""In-process A/B of trainer / plan knobs at the bench shape (64 x 1 x 256 x 256 bf16): one N2NTrainer per configuration
(the knobs are read when the trainer / its plans are created), the configurations timed in interleaved rounds on the same
box and clocks, CUDA events around K graph-replayed steps.  Also checks that every configuration ends with bit-identical
weights (same launches, different stream placement).

usage: python scripts/ab_trainer.py "N2N_OVERLAP=0" "N2N_OVERLAP=1" "N2N_OVERLAP=1 N2N_NO_UPFUSE_TRAIN=1" ...
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from image_denoising_b200 import N2NTrainer, UNet, n2n

B, K, ROUNDS = int(os.environ.get("B", "64")), int(os.environ.get("K", "60")), int(os.environ.get("ROUNDS", "4"))
dev = torch.device("cuda:0")
cfgs = sys.argv[1:] or ["N2N_OVERLAP=0", "N2N_OVERLAP=1"]
gen = torch.Generator(device=dev).manual_seed(100)
batches = []
for _ in range(8):
    clean = torch.rand((B, 1, 256, 256), generator=gen, device=dev)
    batches.append(clean + torch.randn(clean.shape, generator=gen, device=dev) * (25.0 / 255.0))

trainers = []
for cfg in cfgs:
    for kv in cfg.split():
        k, v = kv.split("=")
        os.environ[k] = v
    torch.manual_seed(1234)
    net = UNet(in_nc=1, out_nc=1, n_feature=48).to(dev).set_precision("bf16")
    tr = N2NTrainer(net, lr=3e-4, precision="bf16")
    n2n.operation_seed_counter = 0
    for i in range(4):                       # eager step, capture, replays: plans are created here, under this cfg's knobs
        tr.step(batches[i % 8], 0.02)
    torch.cuda.synchronize()
    trainers.append(tr)

times = [[] for _ in cfgs]
for r in range(ROUNDS):
    for j, tr in enumerate(trainers):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(5):
            tr.step(batches[i % 8], 0.02)
        torch.cuda.synchronize()
        e0.record()
        for i in range(K):
            tr.step(batches[i % 8], 0.02)
        e1.record()
        torch.cuda.synchronize()
        times[j].append(e0.elapsed_time(e1) / K)
for cfg, t in zip(cfgs, times):
    print(f"{cfg:50s} ms/step {' '.join('%.3f' % v for v in t)}  median {statistics.median(t):.4f} min {min(t):.4f}")

# same number of steps on the same batches and selectors -> the weights must agree bit for bit across configurations
finals = []
for tr in trainers:
    torch.manual_seed(1234)
    ref = UNet(in_nc=1, out_nc=1, n_feature=48).to(dev)
    tr.flat_p.copy_(torch.cat([p.detach().reshape(-1) for p in ref.parameters()]))
    for st in (getattr(tr, "flat_m", None), getattr(tr, "flat_v", None)):
        if st is not None:
            st.zero_()
    tr.step_count = 0
    n2n.operation_seed_counter = 0
    for i in range(3):
        l3 = tr.step(batches[i % 8], 0.02)
    torch.cuda.synchronize()
    finals.append((tr.flat_p.clone(), l3.clone()))
same = all(torch.equal(finals[0][0], f[0]) and torch.equal(finals[0][1], f[1]) for f in finals[1:])
print("weights / loss bit-identical across configurations after 3 steps:", same, [float(f[1][0]) for f in finals])
