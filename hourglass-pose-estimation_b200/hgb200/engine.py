"""Inference engine: the stacked-hourglass forward as a static plan of sm_100a kernel launches.

A plan is built once per (batch, height, width, flip) from folded weights: every intermediate lives in
a buffer handed out by a small arena (reused as soon as its last reader is enqueued -- single stream,
so stream order makes reuse safe), every launch is a closure over fixed device pointers, and the
whole sequence (~50 launches per stack) is captured into ONE CUDA graph so the host cost per
forward is a single graph launch.

Dataflow (reference: src/models/hourglass.py:69-90, src/models/modules.py:27-47,80-96):
  stem     im2col(7x7/s2) -> GEMM(+folded bn1, ReLU)
  block    K1: 1x1 GEMM, prologue relu(bn1(x)), epilogue +b (bn2 folded) ReLU
           K2: 3x3 implicit GEMM (9 shifted TMA boxes), epilogue +b (bn3 folded) ReLU
           K3: 1x1 GEMM, epilogue +b +residual [+ nearest-upsampled low-res branch]
                (downsample blocks: residual conv rides along as a second K segment)
  hourglass  low branch first; `up1 + upsample(low3)` is the up1 block's K3 epilogue
  heads    fc (+folded BN, ReLU) -> score (fp32 NCHW heat map) ; remap = ONE merged 256->256 GEMM
"""
from __future__ import annotations

import os
from collections import defaultdict
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import dag, ops
from ._lib import HgError
from .fold import NetWeights, BlockWeights

# The forward's launches can be captured as a dependency DAG across several CUDA streams (hgb200/dag.py) so that the
# `up1` bottleneck of every hourglass level runs beside the lower pyramid.  Measured on B200 (images/s, flip test):
#   batch 1: 453 -> 496 with 4 -> 8 streams;  batch 8: 2341 / 2582 / 2599 for 1 / 4 / 8 streams;  batch 32: 4365 / 4525 / 4571;
#   batch 128: 5190 / 5229 / 5195 / 5172 for 1 / 2 / 4 / 6 streams -- there the persistent GEMM kernels own every SM and
#   nothing is left to overlap.
# So small batches (at most 64 rows through the network, mirrors included) use 8 streams, large ones one chain on one
# stream; HG_INFER_STREAMS forces a count.
_STREAMS_ENV = os.environ.get("HG_INFER_STREAMS")
STREAMS = int(_STREAMS_ENV) if _STREAMS_ENV else 0          # 0 = choose per plan


def streams_for(rows: int) -> int:
    if STREAMS > 0:
        return STREAMS
    return 8 if rows <= 64 else 1


# The 2x2 max-pool of a hourglass level's input is written by the epilogue of the 1x1 GEMM that produces that input
# (hg_conv_desc.pool_out) instead of a separate kernel that re-reads it; HG_NO_POOL_FUSION=1 keeps the separate kernel.
FUSE_POOL = os.environ.get("HG_NO_POOL_FUSION", "0") != "1"
# K2 + K3 of a bottleneck (3x3 -> 1x1 + residual / upsample-add) in one paired-CTA launch wherever the level has enough tiles
# for the CTA pairs: bit-identical to the two kernels, 10-17 % faster on the pair, +4-5 % on the C2 step.
# HG_NO_FUSE_K3=1 keeps the two kernels.
FUSE_K3 = os.environ.get("HG_NO_FUSE_K3") is None
# The 2x2 max-pool of a hourglass level's input comes out of the PROLOGUE of the 1x1 conv that opens the level's `up1`
# bottleneck (hg_conv_desc.pool_in): both read the same tensor (src/models/modules.py:81-83), so that conv is emitted first and
# the pool kernel with its re-read of the 256-channel tensor disappears.  HG_NO_POOL_IN=1 keeps the pool kernel.
POOL_IN = os.environ.get("HG_NO_POOL_IN") is None


def _host_copy(t: torch.Tensor) -> torch.Tensor:
    """CPU copy of a parameter without launching a device kernel: a conv weight that the training engine keeps in
    [co][kh][kw][ci] order (a channels_last view of the flat store) crosses as the dense tensor it is in memory."""
    t = t.detach()
    if t.device.type == "cpu":
        return t
    if t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last):
        return t.permute(0, 2, 3, 1).cpu().permute(0, 3, 1, 2)
    return t.cpu()


def _tensors_of(obj):
    """Every tensor reachable from a NetWeights / BlockWeights tree, in a deterministic order."""
    if torch.is_tensor(obj):
        yield obj
    elif isinstance(obj, (list, tuple)):
        for e in obj:
            yield from _tensors_of(e)
    elif isinstance(obj, (NetWeights, BlockWeights)):
        for k in sorted(vars(obj)):
            yield from _tensors_of(getattr(obj, k))


class _Arena:
    """Buffers recycled in emission order.  With the launch DAG a recycled buffer is a write-after-read edge between
    otherwise independent launches, so with several streams reuse is first-in-first-out and only once `depth`
    buffers of that shape are free."""

    def __init__(self, device, depth: int = 1):
        self.device = device
        self.free = defaultdict(list)
        self.total_bytes = 0
        self.depth = depth
        self.on_get = None

    def get(self, shape, dtype=torch.bfloat16) -> torch.Tensor:
        key = (tuple(shape), dtype)
        if len(self.free[key]) >= self.depth:
            t = self.free[key].pop(0)
        else:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self.total_bytes += t.numel() * t.element_size()
        if self.on_get is not None:
            self.on_get(t)                   # a re-issued buffer forgets what was recorded about its previous content
        return t

    def put(self, t: torch.Tensor):
        self.free[(tuple(t.shape), t.dtype)].append(t)

    # halo-padded activations ([zero row][n][h+1][w+1][c], see hg_conv3x3_halo_bf16): zeroed once at
    # allocation; producers only write the interior, so a recycled buffer's pads are still zero.
    def get_halo(self, n, h, w, c) -> torch.Tensor:
        key = ("halo", n, h, w, c)
        if len(self.free[key]) >= self.depth:
            return self.free[key].pop(0)
        t = ops.halo_padded_buffer(n, h, w, c, self.device)
        self.total_bytes += t.numel() * t.element_size()
        return t

    def put_halo(self, t: torch.Tensor, n, h, w, c):
        self.free[("halo", n, h, w, c)].append(t)


class Plan:
    """A recorded launch sequence with static buffers; `run()` replays it (graph or eager)."""

    def __init__(self, device):
        self.device = device
        self.launches: List[Callable[[], None]] = []
        self.meta: List[dict] = []          # one entry per launch: op class, algorithmic flops / bytes
        self.input: Optional[torch.Tensor] = None
        self.outputs: List[torch.Tensor] = []
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.arena_bytes = 0
        self.streams = 1                    # CUDA streams the graph is captured across (launch DAG when > 1)
        self.heatmap = self.center = self.scale = self.coords = None

    @property
    def num_launches(self) -> int:
        return len(self.launches)

    def run_eager(self):
        for fn in self.launches:
            fn()

    def capture(self):
        # warm up on a side stream (lazy module loading, smem attribute opt-in) while recording what every launch
        # reads and writes, then capture -- as a launch DAG over several streams when STREAMS > 1
        global ops
        self.records = []
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with dag.RECORDING:                 # the rebinding of `ops` below is process-wide: one recording at a time
            rec = dag.Recorder(ops)
            real_ops, ops = ops, rec
            try:
                with torch.cuda.stream(s):
                    for fn in self.launches:
                        rec.calls = []
                        fn()
                        self.records.append(rec.calls)
            finally:
                ops = real_ops
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        ops.check_err_word(self.device)
        if self.streams > 1:
            _, stream_of, waits = self.schedule()
            self.graph = dag.capture(self.launches, stream_of, waits, self.streams, self.device)
            return
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run_eager()
        self.graph = g

    def schedule(self, streams: Optional[int] = None):
        """-> (dag, stream of each launch, cross-stream waits of each launch)."""
        k = streams or self.streams
        d = dag.build(self.records)
        cost = [4e-6 + max(m["flops"] / 6e14, m["bytes"] / 3e12) for m in self.meta]
        stream_of, waits = dag.assign_streams(d, cost, k)
        return d, stream_of, waits

    def profile(self, iters: int = 1):
        """Eager replay with a CUDA-event pair around every launch (on the launching stream).
        Returns per-launch milliseconds averaged over `iters` -- used for the roofline figures."""
        st = torch.cuda.current_stream(self.device)
        total = [0.0] * len(self.launches)
        for _ in range(iters):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.launches) + 1)]
            evs[0].record(st)
            for i, fn in enumerate(self.launches):
                fn()
                evs[i + 1].record(st)
            torch.cuda.synchronize(self.device)
            for i in range(len(self.launches)):
                total[i] += evs[i].elapsed_time(evs[i + 1])
        return [t / iters for t in total]

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run_eager()


class HourglassEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device, depth: int = 4):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise HgError("HourglassEngine needs a CUDA device (no CPU fallback)")
        # BatchNorm folding and the GEMM weight layouts are host work on a CPU copy of the parameters (one D2H memcpy per
        # tensor, no device launches): the device only ever sees the finished bf16 / fp32 operands.
        w = NetWeights({k: _host_copy(v) for k, v in state_dict.items()}, depth)
        self.w = w
        self.num_stacks = w.num_stacks
        self.num_classes = w.num_classes
        self._to_device(w)
        self.plans: Dict[Tuple[int, int, int, bool], Plan] = {}

    def update_weights(self, state_dict: Dict[str, torch.Tensor]):
        """Re-fold new parameter values of the SAME architecture into the existing device operands (in place: device
        pointers, plans and captured graphs stay valid).  What `model.eval()` after a training epoch costs: one D2H copy of
        the parameters, the host fold, one H2D copy per operand -- no plan rebuild."""
        new = NetWeights({k: _host_copy(v) for k, v in state_dict.items()}, self.w.depth)
        old_t, new_t = list(_tensors_of(self.w)), list(_tensors_of(new))
        if len(old_t) != len(new_t) or any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(old_t, new_t)):
            raise HgError("update_weights: the state_dict describes a different architecture")
        for a, b in zip(old_t, new_t):
            a.copy_(b, non_blocking=False)

    def _to_device(self, obj):
        for k, v in list(vars(obj).items()):
            if torch.is_tensor(v):
                setattr(obj, k, v.to(self.device).contiguous())
            elif isinstance(v, (list, tuple)):
                setattr(obj, k, self._conv_list(v))
            elif isinstance(v, BlockWeights):
                self._to_device(v)

    def _conv_list(self, v):
        out = []
        for e in v:
            if torch.is_tensor(e):
                out.append(e.to(self.device).contiguous())
            elif isinstance(e, (list, tuple)):
                out.append(type(e)(self._conv_list(e)))
            elif isinstance(e, BlockWeights):
                self._to_device(e)
                out.append(e)
            else:
                out.append(e)
        return out

    # ------------------------------------------------------------------ plan construction
    def build_plan(self, n: int, h: int, w: int, flip=False, use_graph: bool = True, last_only: bool = False,
                   decode: Optional[dict] = None) -> Plan:
        """flip: False | True (mirrored input) | 'both' (flip test: the batch is doubled internally --
        images [0,n) as given, [n,2n) mirrored -- so the small pyramid levels see twice the rows).
        last_only: emit only the last stack's heat map (inference; the remap GEMM does not need the
        intermediate heat maps).  decode: {'flip_pairs': [...], 'output_size': (w,h)} appends the
        flip-average + get_final_preds_v1 kernels; results in plan.heatmap / plan.coords."""
        if h % 64 or w % 64:
            raise HgError(f"input {h}x{w}: height and width must be multiples of 64 (4 pooling levels on H/4 x W/4)")
        dev = self.device
        W = self.w
        plan = Plan(dev)
        both = (flip == 'both')
        nb = 2 * n if both else n           # rows the network sees
        plan.streams = streams_for(nb) if use_graph else 1
        arena = _Arena(dev, depth=3 if plan.streams > 1 else 1)
        L = plan.launches
        plan.input = torch.zeros((n, 3, h, w), dtype=torch.float32, device=dev)
        # sizes whose 128-pixel tiles are not whole image rows (64x48 maps of 256x192 inputs) store the halo layout pixel by pixel
        ragged_halo = os.environ.get("HG_NO_RAGGED_HALO") is None
        halo_min_w = int(os.environ.get("HG_HALO_MIN_W", "16"))   # levels at least this wide use the halo 3x3 kernel (16x16: 28.5 vs 34.8 us)
        producer: Dict[int, int] = {}       # data_ptr of a fusable 1x1 conv's output -> index of its launch
        pool_targets: Dict[int, torch.Tensor] = {}
        arena.on_get = lambda t: producer.pop(t.data_ptr(), None)

        def conv(x, wt, bias, *, ksize, cout, relu=False, in_scale=None, in_shift=None, residual=None, up_low=None,
                 x2=None, out_f32=None):
            shape = x.shape[:3]
            pixels = shape[0] * shape[1] * shape[2]
            cin_, cin2_ = x.shape[3], (x2.shape[3] if x2 is not None else 0)
            ktot = ksize * ksize * cin_ + cin2_
            act_bytes = pixels * (cin_ + cin2_) * 2 + pixels * cout * (4 if out_f32 is not None else 2)
            if residual is not None:
                act_bytes += pixels * cout * 2
            if up_low is not None:
                act_bytes += pixels * cout * 2 // 4
            plan.meta.append(dict(op=f"conv{ksize}x{ksize}_k{ktot}_n{cout}_{shape[1]}x{shape[2]}"
                                     + ("_pro" if in_scale is not None else "") + ("_res" if residual is not None else "")
                                     + ("_up" if up_low is not None else ""),
                                  kind="conv", flops=2.0 * pixels * ktot * cout, bytes=act_bytes + wt.numel() * 2))
            if out_f32 is not None:
                L.append(lambda: ops.conv_nhwc(x, wt, bias, ksize=ksize, cout=cout, relu=relu, heads=True,
                                               out_nchw_f32=out_f32))
                return out_f32
            out = arena.get((*shape, cout))
            idx = len(L)
            # a later pool(out) may attach its output to THIS launch (pool_targets[idx], looked up when the launch runs)
            if (FUSE_POOL and ksize == 1 and in_scale is None and ops.conv_pool_fusable(shape[1], shape[2], cout, ktot)):
                producer[out.data_ptr()] = idx
            L.append(lambda: ops.conv_nhwc(x, wt, bias, ksize=ksize, cout=cout, relu=relu, in_scale=in_scale,
                                           in_shift=in_shift, residual=residual, up_low=up_low, x2=x2, out=out,
                                           pool_out=pool_targets.get(idx)))
            return out

        def halo_path(bw: BlockWeights, x) -> bool:
            bn_, bh_, bw_, _ = x.shape
            return (not bw.depthwise and bw_ >= halo_min_w and bw_ <= 253 and bw.planes in (64, 128) and x.shape[3] <= 512
                    and (ragged_halo or (bw_ <= 128 and 128 % bw_ == 0 and (bh_ * bw_) % 128 == 0)))

        def block_head(bw: BlockWeights, x, pool_in=None):
            """The bottleneck's opening 1x1 (bn1 + ReLU prologue) into the halo-padded layout; optionally the 2x2 max-pool
            of x as the prologue's second output."""
            bn_, bh_, bw_, _ = x.shape
            a2h = arena.get_halo(bn_, bh_, bw_, bw.planes)
            pixels = bn_ * bh_ * bw_
            plan.meta.append(dict(op=f"conv1x1_k{x.shape[3]}_n{bw.planes}_{bh_}x{bw_}_pro_halo" + ("_poolin" if pool_in is not None else ""),
                                  kind="conv", flops=2.0 * pixels * x.shape[3] * bw.planes,
                                  bytes=pixels * (x.shape[3] + bw.planes) * 2 + bw.w1.numel() * 2
                                  + (pool_in.numel() * 2 if pool_in is not None else 0)))
            L.append(lambda: ops.conv_nhwc(x, bw.w1, bw.b1, ksize=1, cout=bw.planes, relu=True, in_scale=bw.s1,
                                           in_shift=bw.t1, out_halo=a2h, pool_in=pool_in))
            return a2h

        def block(bw: BlockWeights, x, up_low=None, head=None):
            bn_, bh_, bw_, _ = x.shape
            if bw.depthwise:
                # mobile=True: K2 is a depthwise 3x3 stencil on CUDA cores (9 MAC/element, HBM-bound)
                a2 = conv(x, bw.w1, bw.b1, ksize=1, cout=bw.planes, relu=True, in_scale=bw.s1, in_shift=bw.t1)
                a3 = arena.get((bn_, bh_, bw_, bw.planes))
                plan.meta.append(dict(op=f"dwconv3x3_c{bw.planes}_{bh_}x{bw_}", kind="bw",
                                      flops=2.0 * bn_ * bh_ * bw_ * 9 * bw.planes, bytes=a2.numel() * 4))
                L.append(lambda: ops.dwconv3x3(a2, bw.w2, bw.b2, relu=True, out=a3))
                arena.put(a2)
            elif halo_path(bw, x):
                # K1 writes the halo-padded layout; K2 reads every input pixel once (hg_conv3x3.cu)
                a2h = head if head is not None else block_head(bw, x)
                pixels = bn_ * bh_ * bw_
                if (FUSE_K3 and bw.planes == 128 and not bw.downsample and bw.cout == 256
                        and ops.conv3x3_k3_fusable(bn_, bh_, bw_)):
                    # K2 + K3 in one launch on CTA pairs: the 3x3's result never leaves the SM (hg_conv3x3_k3_fused_bf16)
                    out = arena.get((bn_, bh_, bw_, bw.cout))
                    plan.meta.append(dict(op=f"conv3x3h_k3_fused_{bh_}x{bw_}" + ("_up" if up_low is not None else ""), kind="conv",
                                          flops=2.0 * pixels * (9 * 128 * 128 + 128 * 256),
                                          bytes=pixels * (128 + 256 + 256) * 2 + (pixels * 256 * 2 // 4 if up_low is not None else 0)
                                          + bw.w2.numel() * 2 + bw.w3.numel() * 2))
                    L.append(lambda: ops.conv3x3_k3_fused(a2h, bw.w2, bw.b2, bw.w3, bw.b3, n=bn_, h=bh_, w=bw_, residual=x,
                                                          up_low=up_low, out=out))
                    arena.put_halo(a2h, bn_, bh_, bw_, bw.planes)
                    return out
                a3 = arena.get((bn_, bh_, bw_, bw.planes))
                plan.meta.append(dict(op=f"conv3x3h_k{9 * bw.planes}_n{bw.planes}_{bh_}x{bw_}", kind="conv",
                                      flops=2.0 * pixels * 9 * bw.planes * bw.planes,
                                      bytes=pixels * bw.planes * 4 + bw.w2.numel() * 2))
                L.append(lambda: ops.conv3x3_halo(a2h, bw.w2, bw.b2, n=bn_, h=bh_, w=bw_, cin=bw.planes,
                                                  cout=bw.planes, relu=True, out=a3))
                arena.put_halo(a2h, bn_, bh_, bw_, bw.planes)
            else:
                a2 = conv(x, bw.w1, bw.b1, ksize=1, cout=bw.planes, relu=True, in_scale=bw.s1, in_shift=bw.t1)
                a3 = conv(a2, bw.w2, bw.b2, ksize=3, cout=bw.planes, relu=True)
                arena.put(a2)
            if bw.downsample:
                out = conv(a3, bw.w3, bw.b3, ksize=1, cout=bw.cout, x2=x, up_low=up_low)
            else:
                out = conv(a3, bw.w3, bw.b3, ksize=1, cout=bw.cout, residual=x, up_low=up_low)
            arena.put(a3)
            return out

        def chain(blocks, x, up_low=None, keep_input=True, head=None):
            """nn.Sequential of bottlenecks; `up_low` joins the LAST block's epilogue; `head`: the first block's opening
            1x1 has already been emitted (block_head)."""
            cur = x
            for i, bw in enumerate(blocks):
                nxt = block(bw, cur, up_low if i == len(blocks) - 1 else None, head if i == 0 else None)
                if cur is not x or not keep_input:
                    arena.put(cur)
                cur = nxt
            return cur

        def pool(x):
            nb, hh, ww, c = x.shape
            out = arena.get((nb, hh // 2, ww // 2, c))
            idx = producer.pop(x.data_ptr(), None)
            if idx is not None:
                # the 1x1 GEMM that produced x writes the pooled tensor from its own epilogue: no pool launch, and the
                # full re-read of x disappears
                pool_targets[idx] = out
                plan.meta[idx]["op"] += "_pool"
                plan.meta[idx]["bytes"] += out.numel() * 2
                return out
            plan.meta.append(dict(op=f"maxpool_{hh}x{ww}_c{c}", kind="bw", flops=0.0, bytes=x.numel() * 2 * 1.25))
            L.append(lambda: ops.maxpool2x2(x, out))
            return out

        def hourglass(levels, d, x, cat=None):
            """Hourglass._hour_glass_forward(n=d+1, x); x stays owned by the caller."""
            head = None
            up_first = levels[d][0][0]
            nb_, hh_, ww_, c_ = x.shape
            if (POOL_IN and cat is None and x.data_ptr() not in producer and halo_path(up_first, x) and up_first.planes == 128
                    and not up_first.downsample and ops.conv_pool_in_fusable(nb_, hh_, ww_, up_first.planes)):
                # up1's opening 1x1 first: its prologue warps hold every tile of x in shared memory and write the pooled
                # tensor on the way; its halo-padded result waits for the lower pyramid (up1's tail takes low3 along)
                p = arena.get((nb_, hh_ // 2, ww_ // 2, c_))
                head = block_head(up_first, x, pool_in=p)
            else:
                p = pool(x)
            low1 = chain(levels[d][1], p, keep_input=False)
            if d > 0:
                low2 = hourglass(levels, d - 1, low1, cat)
                arena.put(low1)
            else:
                low2 = chain(levels[0][3], low1, keep_input=False)
            low3 = chain(levels[d][2], low2, keep_input=False)
            if cat is None:
                out = chain(levels[d][0], x, up_low=low3, head=head)
                arena.put(low3)
                return out
            # skip_mode='concat': grouped 1x1 over cat([up1, upsample(low3)]) as two zero-padded GEMMs (fold.py)
            wa, ba, wb, bb = cat
            up1 = chain(levels[d][0], x)
            t = conv(low3, wb, bb, ksize=1, cout=wb.shape[0])
            arena.put(low3)
            out = conv(up1, wa, ba, ksize=1, cout=wa.shape[0], up_low=t)
            arena.put(up1)
            arena.put(t)
            return out

        # ---- stem: pack to NHWC4 (+mirror for the flip half) -> windowed implicit GEMM, no im2col matrix
        x_in = plan.input
        packed = ops.stem_packed_buffer(nb, h, w, dev)           # outside the arena: its zero padding must persist
        plan.keep = [packed]
        pack_bytes = n * 3 * h * w * 4 + n * h * w * 8
        if both:
            # both orientations of the flip test from ONE read of the fp32 image
            L.append(lambda a=packed: ops.stem_pack(x_in, a, flip_w="both"))
            plan.meta.append(dict(op="stem_pack_both", kind="bw", flops=0.0, bytes=pack_bytes + n * h * w * 8))
        else:
            L.append(lambda a=packed: ops.stem_pack(x_in, a, flip_w=bool(flip)))
            plan.meta.append(dict(op="stem_pack", kind="bw", flops=0.0, bytes=pack_bytes))
        s0 = arena.get((nb, h // 2, w // 2, 64))
        L.append(lambda a=packed, o=s0: ops.stem_conv(a, W.stem_w_win, W.stem_b, out=o))
        plan.meta.append(dict(op=f"stem_conv7x7s2_{h}x{w}", kind="conv", flops=2.0 * nb * (h // 2) * (w // 2) * 147 * 64,
                              bytes=nb * h * w * 8 + nb * (h // 2) * (w // 2) * 128))
        l1 = chain(W.layer1, s0, keep_input=False)
        p1 = pool(l1)
        arena.put(l1)
        l2 = chain(W.layer2, p1, keep_input=False)
        x = chain(W.layer3, l2, keep_input=False)
        # ---- stacks
        hm_h, hm_w = h // 4, w // 4
        for i in range(self.num_stacks):
            y = hourglass(W.hg[i], W.depth - 1, x, W.concat[i] if W.concat else None)
            y = chain(W.res[i], y, keep_input=False)
            fw, fb = W.fc[i]
            y2 = conv(y, fw, fb, ksize=1, cout=256, relu=True)
            arena.put(y)
            if not last_only or i == self.num_stacks - 1:
                out_i = torch.empty((nb, self.num_classes, hm_h, hm_w), dtype=torch.float32, device=dev)
                sw, sb = W.score[i]
                conv(y2, sw, sb, ksize=1, cout=self.num_classes, out_f32=out_i)
                plan.outputs.append(out_i)
            if i < self.num_stacks - 1:
                mw, mb = W.remap[i]
                x_next = conv(y2, mw, mb, ksize=1, cout=256, residual=x)
                arena.put(x)
                x = x_next
            arena.put(y2)
        if decode is not None:
            hm = plan.outputs[-1]
            if both:
                from .flip import flip_perm_tensor
                perm = flip_perm_tensor(self.num_classes, decode.get('flip_pairs') or [], dev)
                avg = torch.empty((n, self.num_classes, hm_h, hm_w), dtype=torch.float32, device=dev)
                h0, h1 = hm[:n], hm[n:]
                L.append(lambda: ops.flip_average(h0, h1, perm, out=avg))
                plan.meta.append(dict(op="flip_average", kind="bw", flops=0.0, bytes=avg.numel() * 4 * 3))
                hm = avg
            plan.heatmap = hm
            plan.center = torch.zeros((n, 2), dtype=torch.float64, device=dev)
            plan.scale = torch.ones((n, 2), dtype=torch.float64, device=dev)
            plan.coords = torch.empty((n, self.num_classes, 2), dtype=torch.float64, device=dev)
            osz = decode.get('output_size') or (hm_w, hm_h)
            L.append(lambda: ops.decode_final_preds_into(hm, plan.center, plan.scale, plan.coords, osz))
            plan.meta.append(dict(op="decode_final_preds", kind="bw", flops=0.0, bytes=hm.numel() * 4))
        plan.arena_bytes = arena.total_bytes
        if use_graph:
            plan.capture()
        return plan

    def plan_for(self, n, h, w, flip=False, use_graph=True) -> Plan:
        key = (n, h, w, bool(flip), bool(use_graph))
        p = self.plans.get(key)
        if p is None:
            p = self.build_plan(n, h, w, flip, use_graph)
            self.plans[key] = p
        return p

    # ------------------------------------------------------------------ public
    def forward(self, x: torch.Tensor, flip: bool = False, use_graph: bool = True, clone: bool = True):
        """x: fp32 NCHW [n,3,h,w] on this device -> list of num_stacks fp32 NCHW heat maps [n,J,h/4,w/4].

        With clone=False the returned tensors are the plan's static output buffers (overwritten by the
        next forward of the same shape)."""
        if x.dim() != 4 or x.shape[1] != 3:
            raise HgError(f"expected [n,3,h,w], got {tuple(x.shape)}")
        n, _, h, w = x.shape
        if n == 0:      # an empty batch gives empty heat maps, as the reference's modules do
            return [torch.zeros((0, self.num_classes, h // 4, w // 4), dtype=torch.float32, device=self.device)
                    for _ in range(self.num_stacks)]
        plan = self.plan_for(n, h, w, flip, use_graph)
        plan.input.copy_(x, non_blocking=True)
        plan.run()
        return [o.clone() for o in plan.outputs] if clone else list(plan.outputs)
