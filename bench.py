#!/usr/bin/env python
"""bench.py — N2N train patches/s (256^2) on N B200s (BASELINE.json metric, config C3).

One "step" = one Neighbor2Neighbor training iteration (training_script.md:128-156) on a batch of
64 x 1 x 256 x 256 synthetic Gaussian-noised (sigma 25) patches per GPU: fused sub-sampler ->
no-grad full-resolution UNet forward -> half-resolution forward + backward -> fused loss ->
(NCCL all-reduce of the 5 MB gradient when N > 1) -> fused Adam.  Nothing is skipped.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by the driver through torch.distributed.run (one rank per GPU, NCCL).
`--impl reference` times the reference's own algorithm on the host CPU cores (the oracle port of
the pure-Python/PyTorch reference) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "n2n_train_patches_per_s_256x256"
UNIT = "patches/s"
BATCH_PER_GPU = 64
PATCH = 256
NF = 48
# algorithmic FLOPs per 256^2 patch: no-grad fwd 38.573 + fwd@128^2 9.643 + bwd 2*9.643 - 0.014
GFLOP_PER_PATCH = 67.49
# of which the tap-GEMM kernels (conv/deconv forward + input gradients) and the weight-gradient kernel; the
# fused head backward (class "tap-GEMM") also produces the nin_a / nin_b weight gradients:
# 2 layers x 2 x 128^2 px x 96 x 96 = 0.604 GFLOP per patch move from the second line to the first
GFLOP_HEAD_WGRAD_PER_PATCH = 2 * 2 * 128 * 128 * 96 * 96 / 1e9
GFLOP_TAPGEMM_PER_PATCH = 38.573 + 9.643 + (9.643 - 0.014) + GFLOP_HEAD_WGRAD_PER_PATCH
GFLOP_WGRAD_PER_PATCH = 9.643 - GFLOP_HEAD_WGRAD_PER_PATCH


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16=p.get("bf16_tflops_sustained", 1400.0), hbm=p.get("hbm_gbs", 6650.0), src="measured (sustained)")
    return dict(bf16=1400.0, hbm=6650.0, src="fallback")


def _traffic_note():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    `ncu --set full` capture (profiles/roofline_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(path))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region (NVML; falls back to nvidia-smi)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []          # (sm_mhz, max_mhz, reasons bitmask or None, [nvidia-smi reason strings])
        self.stop_flag = False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.samples.append((sm, self.max_sm, mask, None))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            f = [x.strip() for x in out.split(",")]
            self.samples.append((float(f[0]), float(f[1]), None, f[2:6]))

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.02 if self.nvml is not None else 0.2)

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # NVML clocks-event-reason bits: SW power cap 0x4, HW slowdown 0x8, SW thermal 0x20, HW thermal 0x40
        bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            sm.append(s[0]); mx = max(mx, s[1])
            if s[2] is not None:
                for n, b in bits.items():
                    if s[2] & b:
                        reasons.add(n)
            elif s[3]:
                for n, v in zip(names, s[3]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_reference_steps(steps: int, warmup: int, batch: int = 4):
    """The reference algorithm on the host CPU (oracle port, fp32, all torch threads):
    one N2N step + Adam on `batch` 1x256x256 patches (BASELINE.json configs[0])."""
    import torch
    from oracle import n2n_oracle as O
    torch.manual_seed(0)
    p = O.unet_init(1, 1, NF, 0)
    m = {k: torch.zeros_like(v).numpy() for k, v in p.items()}
    v = {k: torch.zeros_like(v).numpy() for k, v in p.items()}
    clean = torch.rand(batch, 1, PATCH, PATCH)
    noisy = clean + torch.randn(clean.shape) * (25.0 / 255.0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        rd = O.draw_rd_idx(batch, PATCH, PATCH, it + 1)
        m1, m2 = O.masks_from_rd_idx(rd)
        _, _, _, grads, _, _ = O.n2n_step_grads(p, noisy, m1, m2, 0.02)
        for k in p:
            O.adam_update(p[k].numpy(), grads[k].numpy(), m[k], v[k], it + 1, 3e-4)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return batch / (sum(times) / len(times)), torch.get_num_threads(), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warmup = max(1, min(args.warmup, 2))
    value, cores, sec = cpu_reference_steps(steps, warmup, batch=4)
    sample = f"{steps} timed N2N steps (+{warmup} warm-up) on batch 4x1x256x256 fp32, oracle port on torch CPU"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "n2n_train_unet48_b64_1x256x256 (BASELINE configs[2])", "batch_per_gpu": BATCH_PER_GPU,
                   "global_batch": BATCH_PER_GPU * max(args.gpus, 1), "patch": PATCH, "n_feature": NF,
                   "parallelism": "host CPU, all torch threads",
                   "sample": "each timed step = the same N2N iteration on a bounded batch of 4 patches (rank 0 only)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def inference_704(dev, precision, world, rank, dist, total_images=128, per_launch=8, reps=2):
    """BASELINE configs[3]: evaluation.py semantics on 128 synthetic 704x704 grayscale images sharded over
    the ranks (no collective): pinned uint8 H2D -> /255 -> whole-image UNet forward -> clamp/quantise ->
    PSNR/SSIM kernel -> D2H of the metrics.  Returns whole-job images/s (device timing, max over ranks)."""
    import torch
    from image_denoising_b200 import UNet, ops
    torch.manual_seed(4321)
    net = UNet(in_nc=1, out_nc=1, n_feature=NF).to(dev).set_precision(precision)
    mine = total_images // world
    g = torch.Generator().manual_seed(2025 + rank)
    clean = (torch.rand((per_launch, 704, 704), generator=g) * 255).to(torch.uint8)
    noisy = (clean.float() + torch.randn(clean.shape, generator=g) * 25.0).clamp(0, 255).to(torch.uint8)
    clean_h, noisy_h = clean.pin_memory(), noisy.pin_memory()
    res_h = torch.empty((per_launch, 2), dtype=torch.float64).pin_memory()
    clean_d = torch.empty_like(clean, device=dev); noisy_d = torch.empty_like(noisy, device=dev)

    def one_pass():
        for _ in range(0, mine, per_launch):
            noisy_d.copy_(noisy_h, non_blocking=True); clean_d.copy_(clean_h, non_blocking=True)
            with torch.no_grad():
                pred = net((noisy_d.float() / 255.0).unsqueeze(1))
            q = ops.quantize_u8(pred, 0.5).squeeze(1)
            res_h.copy_(ops.psnr_ssim_u8(q, clean_d), non_blocking=True)

    one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ips = world * mine * reps / (ms / 1e3)
    return {"metric": "inference_images_per_s_704x704_whole_image", "value": ips, "unit": "images/s",
            "images": world * mine, "per_launch": per_launch, "psnr_first": float(res_h[0, 0]),
            "gflop_per_image": 291.71, "tflops": ips * 291.71 / 1e3,
            "note": "e2e: pinned uint8 H2D + forward + quantise + PSNR/SSIM kernel + D2H of metrics; random-init weights"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from image_denoising_b200 import N2NTrainer, UNet, _ext, n2n

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (image_denoising_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout during the first collective; the contract is ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    L = _ext.lib()
    assert L.n2n_device_ok() == 1, "libn2n_b200 needs an sm_100 device"

    torch.manual_seed(1234)                     # same weights on every rank (then broadcast anyway)
    net = UNet(in_nc=1, out_nc=1, n_feature=NF).to(dev).set_precision(args.precision)
    trainer = N2NTrainer(net, lr=3e-4, precision=args.precision)
    B = args.batch
    # synthetic Gaussian-noised data, 8 distinct device-resident batches (134 MB) rotated per step
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    nbuf = 8
    batches = []
    for _ in range(nbuf):
        clean = torch.rand((B, 1, PATCH, PATCH), generator=gen, device=dev)
        batches.append(clean + torch.randn(clean.shape, generator=gen, device=dev) * (25.0 / 255.0))
    lam = 1 / 100 * 2.0

    def one_step(i):
        return trainer.step(batches[i % nbuf], lam)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i)
    barrier()
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss3 = one_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if rank == 0:
        sampler.stop_flag = True
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = int(trainer.last_launches) * args.steps
    final_loss = float(loss3[0].item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- end-to-end through the public API with HOST buffers (pinned H2D of the batch + D2H of the loss) ----
    # The batch is staged the way a training loop feeds this trainer (image_denoising_b200.prefetch): two device
    # buffers and a copy stream, so the copy of batch i+1 overlaps step i; every step still moves its own
    # 16.8 MB from pinned host memory inside the timed region and the loss is read back every step.
    from image_denoising_b200.prefetch import DevicePrefetcher
    host = [b.cpu().pin_memory() for b in batches[:4]]
    pf = DevicePrefetcher(batches[0])
    loss_host = torch.empty(3, dtype=torch.float32).pin_memory()
    barrier()
    t0 = time.perf_counter()
    pf.put(host[0])
    for i in range(args.steps):
        dbuf = pf.get()
        if i + 1 < args.steps:
            pf.put(host[(i + 1) % len(host)])
        l3 = trainer.step(dbuf, lam)
        pf.release()
        loss_host.copy_(l3, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the loss every step (train.py:364)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * args.steps / e2e_s

    # ---- roofline leg: per-launch CUDA events around the GEMM kernels, same steps, same stream ----
    # every rank runs the same eager steps (they contain the gradient all-reduce); rank 0 reports
    roof = None
    psteps = max(1, min(3, args.steps))
    L.n2n_profile_begin()
    for i in range(psteps):
        one_step(i)
    out = (ctypes.c_double * 6)()
    _ext.check(L.n2n_profile_end(out))
    barrier()
    if rank == 0:
        pk = _peaks()
        tap_ms, tap_flops_exec, tap_n, wg_ms, wg_flops_exec, wg_n = [float(x) for x in out]
        alg_flops = GFLOP_TAPGEMM_PER_PATCH * 1e9 * B * psteps
        achieved = alg_flops / (tap_ms / 1e3) / 1e12 if tap_ms > 0 else 0.0
        roof = {"bound": "tensor",
                "kernel": "slabgemm_umma_kernel (+ head_chain_umma / head_bwd_umma, tapgemm_umma for the shapes the slab engine "
                          "declines): conv/deconv forward + input gradient, %d launches/step" % round(tap_n / psteps),
                "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": achieved / pk["bf16"],
                "traffic": (_traffic_note() or {}).get("bytes_per_launch"), "traffic_detail": _traffic_note(),
                "peak_source": pk["src"],
                "share_of_step": tap_ms / psteps / (ms / args.steps),
                "executed_tflops_incl_padding": tap_flops_exec / (tap_ms / 1e3) / 1e12 if tap_ms > 0 else 0.0,
                "wgrad_kernel": {"achieved": (GFLOP_WGRAD_PER_PATCH * 1e9 * B * psteps) / (wg_ms / 1e3) / 1e12 if wg_ms > 0 else 0.0,
                                 "unit": "TFLOP/s", "ms_per_step": wg_ms / psteps, "launches_per_step": round(wg_n / psteps)},
                "ms_per_step": tap_ms / psteps}

    infer = None
    if not args.no_inference:
        del trainer
        torch.cuda.empty_cache()
        infer = inference_704(dev, args.precision, world, rank, dist)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sec = cpu_reference_steps(2, 1, batch=4)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "2 timed N2N steps (+1 warm-up) on batch 4x1x256x256 fp32 (BASELINE configs[0]), oracle port on torch CPU"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "n2n_train_unet48_b64_1x256x256 (BASELINE configs[2])", "batch_per_gpu": B,
                       "global_batch": B * world, "patch": PATCH, "n_feature": NF, "parallelism": f"dp{world}",
                       "l2": "per-step working set (~6 GB of activations) >> 126 MB L2; inputs rotate over 8 batches (134 MB)",
                       "tflops_per_step_algorithmic": GFLOP_PER_PATCH * B / 1e3},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * PATCH * PATCH * 4, "d2h_bytes_per_step": 12},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
            "step_tflops": GFLOP_PER_PATCH * B * 1e9 / (ms / args.steps / 1e3) / 1e12,
            "inference_704": infer,
            "final_loss": final_loss,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
