"""N2NTrainer — the fused Neighbor2Neighbor training step (training_script.md:128-156)
driven entirely through libn2n_b200: fused sub-sampler, no-grad full-resolution UNet forward,
half-resolution forward + backward, fused loss, (optional) NCCL gradient all-reduce and fused
multi-tensor Adam.  All buffers are allocated once; a step issues no allocations and no host
synchronisation.  Data parallel = one process per GPU (torch.distributed, NCCL): each rank
takes its slice of the global batch, the flat fp32 gradient buffer (5 MB) is all-reduced and
the 1/world factor is folded into the Adam kernel.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _ext, dp, n2n, ops
from ._ext import check, lib, ptr, ptr_array, stream_ptr
from .optim import build_adam_tables


class N2NTrainer:
    def __init__(self, network, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, precision=None, process_group=None,
                 buckets=2, use_graph=None, data_parallel=True):
        self.net = network
        self.precision = precision or network.precision
        network.set_precision(self.precision)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self.pg = process_group
        self.world, self.rank = 1, 0
        if data_parallel and (process_group is not None or
                              (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        params = list(network.parameters())
        dev = params[0].device
        _ext.require_cuda(params[0], "N2NTrainer")
        sizes = [p.numel() for p in params]
        # one flat fp32 buffer each for params / grads / exp_avg / exp_avg_sq; the module's
        # parameters become views of the flat parameter buffer (state_dict() is unchanged).
        self.flat_p = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        off = 0
        self.params, self.grads = [], []
        for p, n in zip(params, sizes):
            view = self.flat_p[off:off + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            self.params.append(p)
            self.grads.append(self.flat_g[off:off + n].view(p.shape))
            off += n
        self.table, self.blocks = build_adam_tables([self.flat_p], [self.flat_g], [self.flat_m], [self.flat_v], dev)
        # gradient buckets in reverse-autograd order: the head / full-resolution decoder tensors
        # sit at the END of the state_dict order, so bucket 0 is the tail of the flat buffer.
        self.bucket_slices = dp.bucket_slices(sizes, int(os.environ.get("N2N_BUCKETS", buckets)))
        self._shapes = None
        self.last_launches = 0
        # CUDA-graph replay of the whole iteration (one graph launch instead of ~190 kernel launches);
        # N2N_NO_GRAPH=1 keeps the eager launch sequence (used by the per-launch profiling hooks).
        if use_graph is None:
            use_graph = os.environ.get("N2N_NO_GRAPH", "0") != "1"
        self.use_graph = bool(use_graph)
        # the no-grad full-resolution pass and the training forward are independent: two streams (the launch-bound deep
        # levels of one pass fill the SMs the other leaves idle; measured 5.27 -> 5.14 ms/step).  N2N_OVERLAP=0: one stream
        self.overlap_forwards = os.environ.get("N2N_OVERLAP", "1") == "1"
        self._graph = None
        self._eager_steps = 0
        self._rd_global = None          # world > 1: the global-batch selector is drawn on its own stream (see _draw_global_selector)
        self._rd_borrowed = False
        if self.world > 1:
            dp.broadcast_params(self.flat_p, 0, self.pg)

    # ------------------------------------------------------------------ buffers / plans
    def _prepare(self, noisy):
        n, c, h, w = noisy.shape
        if self._shapes == (n, c, h, w):
            return
        net, dev = self.net, noisy.device
        dt = _ext.dtype_tag(self.precision)
        self._destroy_plans()                       # a shape change rebuilds both plans: release the old ones
        self._shapes = (n, c, h, w)
        self.plan_full = ctypes.c_void_p()
        self.plan_half = ctypes.c_void_p()
        check(lib().n2n_unet_plan_create(ctypes.byref(self.plan_full), net.in_nc, net.out_nc, net.n_feature, n, h, w, dt, 0))
        check(lib().n2n_unet_plan_create(ctypes.byref(self.plan_half), net.in_nc, net.out_nc, net.n_feature, n,
                                         h // 2, w // 2, dt, 1))
        self.ws_full = torch.empty(lib().n2n_unet_workspace_bytes(self.plan_full), dtype=torch.uint8, device=dev)
        self.ws_half = torch.empty(lib().n2n_unet_workspace_bytes(self.plan_half), dtype=torch.uint8, device=dev)
        # the half-resolution pass runs on the same weights as the full-resolution one: borrow its packed forward weights
        # instead of repacking them (two streams: the full plan packs BEFORE the fork, n2n_unet_pack_weights)
        self.shared_weights = False
        if os.environ.get("N2N_NO_SHARE", "0") != "1":
            rc = lib().n2n_unet_share_weights(self.plan_half, self.plan_full, ptr(self.ws_full))
            if rc < 0:
                check(rc)
            self.shared_weights = rc == 0
        f32 = dict(dtype=torch.float32, device=dev)
        self.den = torch.empty((n, net.out_nc, h, w), **f32)
        half = (n, c, h // 2, w // 2)
        self.sub1 = torch.empty(half, **f32); self.sub2 = torch.empty(half, **f32)
        self.den1 = torch.empty((n, net.out_nc, h // 2, w // 2), **f32); self.den2 = torch.empty_like(self.den1)
        self.out = torch.empty_like(self.den1)
        self.dout = torch.empty_like(self.den1)
        self.loss3 = torch.zeros(3, **f32)
        self.packed = torch.empty(n * (h // 2) * (w // 2), dtype=torch.uint8, device=dev)
        self.loss_ws = torch.zeros(lib().n2n_loss_workspace_bytes(0), dtype=torch.uint8, device=dev)
        self.param_ptrs = ptr_array(self.params)
        self.grad_ptrs = ptr_array(self.grads)
        self.side_stream = torch.cuda.Stream(device=dev)
        self.ev_fork = torch.cuda.Event(); self.ev_join = torch.cuda.Event()

    def _destroy_plans(self):
        for name in ("plan_full", "plan_half"):
            h = getattr(self, name, None)
            if h is not None and h.value:
                lib().n2n_unet_plan_destroy(h)
            setattr(self, name, None)

    def __del__(self):
        try:
            self._graph = None
            self._destroy_plans()
        except Exception:
            pass

    # ------------------------------------------------------------------ one iteration
    def _launch_sequence(self, noisy, rd_idx, lam, lr, dev_scalars):
        """The launch sequence of one iteration on the current stream.  With ``dev_scalars`` the
        per-step scalars (Lambda, Adam step size / bias correction) are read from device memory, so
        the same sequence can be replayed as a CUDA graph."""
        L, st = lib(), stream_ptr()
        n, c, h, w = noisy.shape
        check(L.n2n_mask_pair_from_rdidx(ptr(rd_idx), rd_idx.numel(), None, None, ptr(self.packed), st))
        check(L.n2n_subsample_pair(ptr(noisy), None, None, ptr(self.packed), ptr(self.sub1), ptr(self.sub2), n, c, h, w, 4, st))
        # The no-grad full-resolution pass and the half-resolution training forward are independent
        # (training_script.md:139-146): run the latter on a side stream so that its launch-bound deep
        # levels fill the SMs the other pass leaves idle (and vice versa); joined before the loss.
        side = self.side_stream if (self.overlap_forwards and L.n2n_profile_active() != 1) else None   # per-launch timing: one stream
        if side is not None:
            if self.shared_weights:
                check(L.n2n_unet_pack_weights(self.plan_full, self.param_ptrs, ptr(self.ws_full), st))
            self.ev_fork.record()
            side.wait_event(self.ev_fork)
            with torch.cuda.stream(side):
                check(L.n2n_unet_forward(self.plan_half, self.param_ptrs, ptr(self.sub1), ptr(self.out), ptr(self.ws_half),
                                         stream_ptr()))
                self.ev_join.record()
        check(L.n2n_unet_forward(self.plan_full, self.param_ptrs, ptr(noisy), ptr(self.den), ptr(self.ws_full), st))
        check(L.n2n_subsample_pair(ptr(self.den), None, None, ptr(self.packed), ptr(self.den1), ptr(self.den2),
                                   n, self.net.out_nc, h, w, 4, st))
        if side is not None:
            torch.cuda.current_stream().wait_event(self.ev_join)
        else:
            check(L.n2n_unet_forward(self.plan_half, self.param_ptrs, ptr(self.sub1), ptr(self.out), ptr(self.ws_half), st))
        if dev_scalars is None:
            check(L.n2n_loss_n2n_fwdbwd(ptr(self.out), ptr(self.sub2), ptr(self.den1), ptr(self.den2), float(lam), 1.0,
                                        self.out.numel(), ptr(self.loss3), ptr(self.dout), ptr(self.loss_ws), st))
        else:
            check(L.n2n_loss_n2n_fwdbwd_dev(ptr(self.out), ptr(self.sub2), ptr(self.den1), ptr(self.den2), ptr(dev_scalars),
                                            1.0, self.out.numel(), ptr(self.loss3), ptr(self.dout), ptr(self.loss_ws), st))
        check(L.n2n_unet_backward(self.plan_half, self.param_ptrs, ptr(self.dout), self.grad_ptrs, None, ptr(self.ws_half), st))
        if self.world > 1:
            dp.allreduce_buckets(self.flat_g, self.bucket_slices, self.pg)
        if dev_scalars is None:
            check(L.n2n_adam_multi(ptr(self.table), 1, ptr(self.blocks), self.blocks.shape[0], float(lr),
                                   float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_count,
                                   1.0 / self.world, st))
        else:
            check(L.n2n_adam_multi_dev(ptr(self.table), 1, ptr(self.blocks), self.blocks.shape[0], ptr(dev_scalars),
                                       float(self.betas[0]), float(self.betas[1]), float(self.eps), 1.0 / self.world, st))

    def _capture(self):
        """Capture the iteration once (static input / selector / scalar buffers)."""
        self.in_static = torch.empty(self._shapes, dtype=torch.float32, device=self.flat_p.device)
        n, c, h, w = self._shapes
        self.rd_static = torch.zeros(n * (h // 2) * (w // 2), dtype=torch.int64, device=self.flat_p.device)
        self.dev_scalars = torch.zeros(4, dtype=torch.float32, device=self.flat_p.device)
        # a capture must not change the optimiser state: snapshot, capture (the captured launches do
        # not execute), nothing to restore
        g = torch.cuda.CUDAGraph()
        launches0 = lib().n2n_launch_count()
        with torch.cuda.graph(g):
            self._launch_sequence(self.in_static, self.rd_static, 0.0, self.lr, self.dev_scalars)
        self.graph_launches = lib().n2n_launch_count() - launches0
        self._graph = g

    def input_buffer(self):
        """The graph's static input tensor (after the first two steps of a shape): a caller that writes its batch here
        (e.g. as the destination of its host-to-device copy) and passes it to ``step`` saves the staging copy."""
        return getattr(self, "in_static", None) if self._graph is not None else None

    def _draw_global_selector(self, noisy):
        """train.py:155-162 for the GLOBAL batch of a W-rank run (same counter seed on every rank; this rank keeps its slice,
        SURVEY.md §8e).  The draw is W times the size of a rank's own selector (67 MB of int64 at W = 8, 64 x 256 x 256 per rank)
        and depends on nothing but the seed counter, so it is issued on its own stream: it runs under the tail of the previous
        step instead of between two steps.  One persistent buffer; the next draw waits for the consumer of this one
        (``_release_global_selector``)."""
        n, c, h, w = noisy.shape
        cells = n * self.world * (h // 2) * (w // 2)
        if self._rd_global is None or self._rd_global.numel() != cells or self._rd_global.device != noisy.device:
            self._rd_global = torch.empty(cells, dtype=torch.int64, device=noisy.device)
            self._rng_stream = torch.cuda.Stream(device=noisy.device)
            self._ev_drawn, self._ev_taken = torch.cuda.Event(), torch.cuda.Event()
            self._rng_stream.wait_stream(torch.cuda.current_stream())       # the buffer's allocation / earlier users
            self._rd_taken = False
        gen = n2n.get_generator(noisy.device)                               # same call order as a 1-process run
        with torch.cuda.stream(self._rng_stream):
            if self._rd_taken:
                self._rng_stream.wait_event(self._ev_taken)
            torch.randint(low=0, high=8, size=(cells,), generator=gen, out=self._rd_global)
            self._ev_drawn.record()
        torch.cuda.current_stream().wait_event(self._ev_drawn)
        self._rd_borrowed = True
        return dp.shard_selector(self._rd_global, self.rank, self.world, n * self.world)

    def _release_global_selector(self):
        """The consumer of the last _draw_global_selector slice has been enqueued on the current stream."""
        if self._rd_borrowed:
            self._ev_taken.record()
            self._rd_taken = True
            self._rd_borrowed = False

    def step(self, noisy, Lambda, rd_idx=None, lr=None):
        """One N2N iteration on this rank's ``noisy`` batch [n,c,h,w] (fp32, CUDA).  Returns the
        device tensor [loss_all, loss1, loss2] (no host sync; the SAME buffer every step — clone it to keep a
        value across steps).  Without ``rd_idx`` the selector is drawn as train.py:155-162 does; with W > 1 ranks
        it is drawn for the GLOBAL batch (identical counter seed on every rank) and this rank keeps its slice, so
        a W-rank run uses the masks a 1-GPU run on the concatenated batch would.  ``rd_idx`` may
        also be passed explicitly (this rank's slice)."""
        noisy = noisy.contiguous()
        shapes_before = self._shapes
        self._prepare(noisy)
        if shapes_before != self._shapes:
            self._graph = None
            self._eager_steps = 0
        L = lib()
        launches0 = L.n2n_launch_count()
        self.step_count += 1
        lr = float(self.lr if lr is None else lr)
        profiling = L.n2n_profile_active() == 1
        replay = self.use_graph and not profiling and self._eager_steps >= 1
        if replay and self._graph is None:
            self._capture()
        if rd_idx is None:
            if self.world > 1:
                rd_idx = self._draw_global_selector(noisy)
            elif replay:
                # train.py:155-162 drawn straight into the graph's selector buffer (no staging copy)
                torch.randint(low=0, high=8, size=(self.rd_static.numel(),), generator=n2n.get_generator(noisy.device),
                              out=self.rd_static)
                rd_idx = self.rd_static
            else:
                rd_idx = n2n.draw_rd_idx(noisy)
        if replay:
            if noisy.data_ptr() != self.in_static.data_ptr():       # callers may fill input_buffer() directly
                self.in_static.copy_(noisy, non_blocking=True)
            if rd_idx.data_ptr() != self.rd_static.data_ptr():
                self.rd_static.copy_(rd_idx, non_blocking=True)
            self._release_global_selector()
            check(L.n2n_set_step_scalars(ptr(self.dev_scalars), float(Lambda), lr, float(self.betas[0]),
                                         float(self.betas[1]), self.step_count, stream_ptr()))
            self._graph.replay()
            self.last_launches = self.graph_launches + 1
            return self.loss3
        self._launch_sequence(noisy, rd_idx, Lambda, lr, None)
        self._release_global_selector()
        self._eager_steps += 1
        self.last_launches = L.n2n_launch_count() - launches0
        return self.loss3
