"""Host -> device staging for the training loop.

The reference feeds the network from a ``DataLoader`` and moves each batch with ``.cuda()`` right before the
step (train.py:343-349): the copy sits on the step's critical path.  ``DevicePrefetcher`` keeps two device
buffers and a copy stream, so batch i+1 crosses PCIe (pinned memory, ``non_blocking``) while step i computes;
the compute stream only waits on the event of the buffer it is about to read, and the copy stream only
overwrites a buffer after the step that read it has been enqueued behind its "free" event."""
from __future__ import annotations

import torch


class DevicePrefetcher:
    def __init__(self, like: torch.Tensor, depth: int = 2):
        assert like.is_cuda and depth >= 2
        self.bufs = [torch.empty_like(like) for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.used = [False] * depth
        self.copy_stream = torch.cuda.Stream(device=like.device)
        self.head = 0          # next slot to fill
        self.tail = 0          # next slot to hand out
        self.bytes_copied = 0

    def put(self, host: torch.Tensor) -> None:
        """Enqueue the H2D copy of one pinned host batch into the next slot (copy stream)."""
        s = self.head
        self.head = (self.head + 1) % len(self.bufs)
        with torch.cuda.stream(self.copy_stream):
            if self.used[s]:
                self.copy_stream.wait_event(self.free[s])       # the step that read this slot is done with it
            self.bufs[s].copy_(host, non_blocking=True)
            self.ready[s].record(self.copy_stream)
        self.bytes_copied += host.numel() * host.element_size()

    def get(self) -> torch.Tensor:
        """The oldest staged batch; the current stream waits for its copy."""
        s = self.tail
        torch.cuda.current_stream().wait_event(self.ready[s])
        return self.bufs[s]

    def release(self) -> None:
        """Call after the consumer of the last get() has been enqueued on the current stream."""
        s = self.tail
        self.tail = (self.tail + 1) % len(self.bufs)
        self.free[s].record(torch.cuda.current_stream())
        self.used[s] = True
