"""Double-buffered host -> device staging for the hot-path inputs (a training input pipeline's prefetch).

The step's inputs (region features, word embeddings, lengths, class ids) arrive in pinned host memory; copying
them on the compute stream exposes the whole transfer (81 MB per rank and step at global batch 2048 on 8 GPUs:
3.4 ms of an 15 ms step).  HostPrefetcher keeps `depth` device-side buffer sets and copies set i+1 on its own stream
while the kernels of step i run; the hand-over in both directions is by CUDA events, nothing synchronises with the
host.  torch is used for streams, events and device memory only.

    pf = HostPrefetcher(device)
    pf.submit(host_tensors)                  # step 0's inputs
    for i in range(steps):
        dev_tensors = pf.acquire()           # compute stream waits for copy i
        if i + 1 < steps:
            pf.submit(next_host_tensors)     # copy i+1 overlaps the kernels of step i
        run_step(dev_tensors)
        pf.release()                         # buffer set i may be overwritten once step i is done
"""
from __future__ import annotations

from collections import deque
from typing import List, Optional, Sequence

import torch


class HostPrefetcher:
    def __init__(self, device, depth: int = 2):
        if not torch.cuda.is_available():
            raise RuntimeError("HostPrefetcher needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device)
        self.depth = depth
        with torch.cuda.device(self.device):
            self.copy_stream = torch.cuda.Stream()
        self.sets: List[Optional[List[torch.Tensor]]] = [None] * depth
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [None] * depth          # recorded on the compute stream when the consumer is done with set k
        self.queue = deque()                # submitted, not yet acquired
        self.in_use = deque()               # acquired, not yet released
        self.next_slot = 0
        self.bytes_copied = 0

    def submit(self, host_tensors: Sequence[torch.Tensor]) -> None:
        k = self.next_slot
        if len(self.queue) + len(self.in_use) >= self.depth:
            raise RuntimeError("HostPrefetcher: every buffer set is in flight (release() before submitting more)")
        self.next_slot = (k + 1) % self.depth
        if self.sets[k] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(self.sets[k], host_tensors)):
            self.sets[k] = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_tensors]
        with torch.cuda.stream(self.copy_stream):
            if self.free[k] is not None:
                self.copy_stream.wait_event(self.free[k])     # the step that last read this set has finished
            for d, h in zip(self.sets[k], host_tensors):
                d.detach().copy_(h, non_blocking=True)        # (a consumer may have set requires_grad on the buffer)
                self.bytes_copied += h.numel() * h.element_size()
            self.ready[k].record(self.copy_stream)
        self.queue.append(k)

    def acquire(self) -> List[torch.Tensor]:
        k = self.queue.popleft()
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        self.in_use.append(k)
        return self.sets[k]

    def release(self) -> None:
        k = self.in_use.popleft()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[k] = ev
