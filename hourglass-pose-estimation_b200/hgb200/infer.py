"""Batched flip-test inference + decode: the serving-side public API of the B200 path.

One CUDA graph per (batch, resolution) does: stem im2col of the batch and of its mirror image ->
8-stack hourglass on the doubled batch -> last-stack heat maps -> flip-average -> per-joint arg-max,
quarter-pixel refine and inverse affine (get_final_preds_v1 for every image).  `infer_host` adds a
double-buffered pinned-host -> device input pipeline so the H2D copy of batch i+1 overlaps the
compute of batch i.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Tuple

import numpy as np
import torch

from . import ops as _ops
from ._lib import HgError
from .engine import HourglassEngine, Plan
from .flip import MPII_FLIP_PAIRS


class FlipTestPipeline:
    def __init__(self, engine: HourglassEngine, batch: int, height: int, width: int,
                 flip_pairs=MPII_FLIP_PAIRS, output_size: Optional[Tuple[int, int]] = None, flip_test: bool = True):
        self.engine = engine
        self.device = engine.device
        self.batch, self.h, self.w = batch, height, width
        self.plan: Plan = engine.build_plan(
            batch, height, width, flip='both' if flip_test else False, use_graph=True, last_only=True,
            decode={'flip_pairs': flip_pairs, 'output_size': output_size or (width // 4, height // 4)})
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._stage = [torch.empty_like(self.plan.input) for _ in range(2)]
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]

    def _check(self, err_host: torch.Tensor):
        """The device error word rides back with every batch's coordinates (no extra synchronisation): a kernel-side
        protocol timeout raises here instead of handing out garbage coordinates."""
        v = int(err_host[0])
        if v != 0:
            _ops.err_word(self.device).zero_()
            raise HgError(f"device-side protocol timeout, code 0x{v & 0xffffffff:x} (role<<8 | barrier)")

    @property
    def launches_per_batch(self) -> int:
        return self.plan.num_launches

    def set_affine(self, center, scale):
        self.plan.center.copy_(torch.as_tensor(np.asarray(center, dtype=np.float64).reshape(self.batch, 2)))
        self.plan.scale.copy_(torch.as_tensor(np.asarray(scale, dtype=np.float64).reshape(self.batch, 2)))

    def infer_device(self, x: torch.Tensor) -> torch.Tensor:
        """x: fp32 NCHW on the device.  Returns the plan's static fp64 [B,J,2] coordinates tensor."""
        self.plan.input.copy_(x, non_blocking=True)
        self.plan.run()
        return self.plan.coords

    def infer_host_u8(self, batches: Iterable[torch.Tensor], mean, std) -> Iterator[np.ndarray]:
        """batches: pinned uint8 NHWC host tensors [B,h,w,3] (cropped frames as the dataset's warpAffine leaves them).
        ToTensor + Normalize(mean, std) runs on the device (hg_normalize_u8_nhwc), so only a quarter of infer_host's
        bytes cross PCIe.  Yields fp64 [B,J,2] numpy arrays."""
        from . import ops
        main = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_stage_u8"):
            self._stage_u8 = [torch.empty((self.batch, self.h, self.w, 3), dtype=torch.uint8, device=self.device)
                              for _ in range(2)]
        out_host = [torch.empty(self.plan.coords.shape, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        err_host = [torch.zeros(1, dtype=torch.int32, pin_memory=True) for _ in range(2)]
        err_dev = _ops.err_word(self.device)
        for s in range(2):
            self._consumed[s].record(main)
        pending = None
        slot = 0
        for xh in batches:
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self._consumed[slot])
                self._stage_u8[slot].copy_(xh, non_blocking=True)
                self._ready[slot].record(self.copy_stream)
            main.wait_event(self._ready[slot])
            ops.normalize_u8(self._stage_u8[slot], mean, std, out=self.plan.input)
            self._consumed[slot].record(main)
            self.plan.run()
            out_host[slot].copy_(self.plan.coords, non_blocking=True)
            err_host[slot].copy_(err_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:
                pending[0].synchronize()
                self._check(pending[2])
                yield pending[1].numpy().copy()
            pending = (done, out_host[slot], err_host[slot])
            slot ^= 1
        if pending is not None:
            pending[0].synchronize()
            self._check(pending[2])
            yield pending[1].numpy().copy()

    def infer_host(self, batches: Iterable[torch.Tensor]) -> Iterator[np.ndarray]:
        """batches: pinned fp32 NCHW host tensors.  Yields fp64 [B,J,2] numpy arrays, one per batch, with
        the next batch's H2D copy in flight while the current one computes."""
        main = torch.cuda.current_stream(self.device)
        it = iter(batches)
        slot = 0
        pending = None
        out_host = [torch.empty(self.plan.coords.shape, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        err_host = [torch.zeros(1, dtype=torch.int32, pin_memory=True) for _ in range(2)]
        err_dev = _ops.err_word(self.device)

        def stage(xh, s):
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self._consumed[s])
                self._stage[s].copy_(xh, non_blocking=True)
                self._ready[s].record(self.copy_stream)

        for s in range(2):
            self._consumed[s].record(main)
        nxt = next(it, None)
        if nxt is not None:
            stage(nxt, slot)
        while nxt is not None:
            cur_slot = slot
            nxt = next(it, None)
            if nxt is not None:
                stage(nxt, cur_slot ^ 1)
            main.wait_event(self._ready[cur_slot])
            self.plan.input.copy_(self._stage[cur_slot], non_blocking=True)
            self._consumed[cur_slot].record(main)
            self.plan.run()
            out_host[cur_slot].copy_(self.plan.coords, non_blocking=True)
            err_host[cur_slot].copy_(err_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:
                pending[0].synchronize()
                self._check(pending[2])
                yield pending[1].numpy().copy()
            pending = (done, out_host[cur_slot], err_host[cur_slot])
            slot ^= 1
        if pending is not None:
            pending[0].synchronize()
            self._check(pending[2])
            yield pending[1].numpy().copy()
