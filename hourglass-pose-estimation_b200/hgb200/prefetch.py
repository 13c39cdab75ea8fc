"""Host <-> device plumbing around the step: the input batch of step i+1 crosses PCIe on a side stream while step i
computes, and the step's scalar result (the loss) is copied back every step but READ one step late, so neither copy nor
the host's wait sits between two steps.  Plumbing only (torch streams / events); no arithmetic happens here."""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence, Tuple

import torch


class DevicePrefetcher:
    """Iterates tuples of device tensors for an iterable of tuples of (ideally pinned) host tensors, `depth` batches
    ahead.  The consumer's stream waits for the copy; the copy stream waits until the consumer has ENQUEUED all work on
    the buffer it is about to overwrite (the consumer signals that by asking for the next batch)."""

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device, depth: int = 2):
        self.device = torch.device(device)
        self.it = iter(batches)
        self.depth = max(2, depth)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots: list = [None] * self.depth
        self.ready = [torch.cuda.Event() for _ in range(self.depth)]
        self.consumed = [torch.cuda.Event() for _ in range(self.depth)]

    def preallocate(self, example: Sequence[torch.Tensor]):
        """Allocate the device staging buffers for batches shaped like `example` now (not inside the first steps)."""
        for s in range(self.depth):
            self.slots[s] = tuple(torch.empty(b.shape, dtype=b.dtype, device=self.device) for b in example)
        return self

    def _stage(self, slot: int) -> bool:
        batch = next(self.it, None)
        if batch is None:
            return False
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])
            cur = self.slots[slot]
            if cur is None or any(c.shape != b.shape or c.dtype != b.dtype for c, b in zip(cur, batch)):
                cur = tuple(torch.empty(b.shape, dtype=b.dtype, device=self.device) for b in batch)
                self.slots[slot] = cur
            for c, b in zip(cur, batch):
                c.copy_(b, non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        return True

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        main = torch.cuda.current_stream(self.device)
        for s in range(self.depth):
            self.consumed[s].record(main)
        pending = []
        for s in range(self.depth - 1):
            if self._stage(s):
                pending.append(s)
        nxt = self.depth - 1
        while pending:
            cur = pending.pop(0)
            if self._stage(nxt):
                pending.append(nxt)
                nxt = (nxt + 1) % self.depth
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self.ready[cur])
            yield self.slots[cur]
            self.consumed[cur].record(torch.cuda.current_stream(self.device))


class LaggedScalar:
    """push(device scalar) copies it to pinned host memory on the current stream and returns the PREVIOUS step's value
    (None on the first call); flush() returns the last one.  The device->host copy happens every step; only the host's
    wait for it is deferred by one step."""

    def __init__(self, device, dtype=torch.float32):
        self.host = [torch.zeros(1, dtype=dtype).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event() for _ in range(2)]
        self.device = torch.device(device)
        self.n = 0

    def push(self, value: torch.Tensor) -> Optional[float]:
        slot = self.n & 1
        prev = self.flush() if self.n > 0 else None
        self.host[slot].copy_(value.reshape(1), non_blocking=True)
        self.events[slot].record(torch.cuda.current_stream(self.device))
        self.n += 1
        return prev

    def flush(self) -> Optional[float]:
        if self.n == 0:
            return None
        slot = (self.n - 1) & 1
        self.events[slot].synchronize()
        return float(self.host[slot][0])
