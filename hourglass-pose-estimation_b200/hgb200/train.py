"""Training engine: one optimisation step of the stacked hourglass (reference: Trainer._train_epoch,
src/runner/trainer.py:82-99) as a static plan of sm_100a kernel launches.

  * Parameters live in ONE flat fp32 buffer (gradients and RMSprop state in two more); the model's
    nn.Parameters are re-pointed to views of it, conv weights in the GEMM-natural [co][kh][kw][ci] order
    (= torch channels_last), so state_dict()/load_state_dict()/torch.optim keep working while the wgrad
    kernels write coalesced and one NCCL all-reduce / one RMSprop launch cover the whole network.
  * Train-mode BatchNorm uses batch statistics, so nothing folds: every BN is a per-channel sum pass
    (hg_colstats_nhwc) plus one apply pass that materialises Z = relu(bn(a)) -- the next GEMM's operand
    AND the wgrad operand kept for backward.
  * Backward per convolution: dgrad = the forward tcgen05 GEMM kernels on transposed (3x3: tap-flipped)
    bf16 weight copies; wgrad = hg_wgrad_bf16 (MN-major split-K GEMM over pixels); BN backward = two
    bandwidth passes, the second also adds the residual-path gradients and writes the halo layout the
    3x3 kernels read.
  * The plan is built once per (batch, height, width): the forward pass records nodes, the backward
    pass is emitted by walking them in reverse with gradient buffers recycled in execution order, and
    the whole step is captured into CUDA graphs.
"""
from __future__ import annotations

import os
from collections import defaultdict
from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn

from . import dag, ops
from ._lib import HgError

RMSPROP_ALPHA = 0.99      # torch.optim.RMSprop defaults (trainer.py:39-41 passes lr, momentum=0, weight_decay=0)
RMSPROP_EPS = 1e-8
BN_EPS = 1e-5
BN_MOMENTUM = 0.1

_ACT = torch.bfloat16     # activation / GEMM-weight storage type
# Data parallel: the gradient all-reduce is issued PER BUCKET INSIDE the step's CUDA graph, on a stream of its own, as soon
# as the last weight-gradient launch of the bucket has run (the hourglass of stack 8 first, ... , the stem last), so the
# exchange over NVLink overlaps the rest of the backward pass.  Opt-in (HG_OVERLAP_AR=1); the default is one
# all-reduce of the whole flat buffer after the graph.  A graph that holds NCCL nodes must be destroyed BEFORE the process
# group is (ncclCommDestroy waits for it): call TrainEngine.release_graphs() first.
OVERLAP_ALLREDUCE = os.environ.get("HG_OVERLAP_AR", "0") == "1"
# The step's launches are captured as a dependency DAG across this many CUDA streams (hgb200/dag.py);
# 1 = one chain on one stream.
STREAMS = int(os.environ.get("HG_TRAIN_STREAMS", "16"))
# ... of which this many are reserved for the leaves of the DAG (weight-gradient GEMMs, bias sums) and run at normal
# priority while the streams carrying the critical chain get CUDA's high priority (0 = no separation).
LEAF_STREAMS = int(os.environ.get("HG_TRAIN_LEAF_STREAMS", "0"))
# HG_BN_STATS_FUSED=1: the batch statistics of a train-mode BatchNorm come out of the epilogue of the GEMM that
# produces its input (hg_conv_desc.stats) for tensors of at most HG_BN_STATS_MAX_PIXELS pixels, instead of a separate
# per-channel sum pass.  OFF by default -- measured on B200 (batch 32, 6 streams): 27.6 ms/step fused (32x32 and
# below; 256 fewer launches) against 27.1 ms with the separate pass, and 29.3 ms when the 64x64 level is fused too
# (there the four epilogue warps become the bottleneck of the HBM-bound 1x1 kernels: +9..47 us per launch against the
# 15-20 us pass they replace).  With the launch DAG the small statistics passes already hide behind other work.
FUSED_STATS = os.environ.get("HG_BN_STATS_FUSED", "0") == "1"
# The 2x2 max-pool of a hourglass level's input comes out of the epilogue of the 1x1 GEMM that produces that input
# (hg_conv_desc.pool_out); HG_NO_POOL_FUSION=1 keeps the separate pool kernel.
FUSE_POOL = os.environ.get("HG_NO_POOL_FUSION", "0") != "1"
FUSED_STATS_MAX_PIXELS = int(os.environ.get("HG_BN_STATS_MAX_PIXELS", "32768"))
# HG_DETERMINISTIC=1: every cross-CTA reduction of the step runs in a fixed order -- BatchNorm statistics, BatchNorm backward
# sums and bias gradients through per-launch scratch slots added by the last CTA to arrive (hg_colstats_nhwc / hg_bn_bwd_reduce
# with `scratch`), weight gradients without the split over the pixels (hg_wgrad_bf16 max_ctas = 1) -- so two runs of the step,
# on one stream or across the launch DAG's streams, with or without programmatic dependent launch, give BIT-IDENTICAL
# heat maps, gradients and parameters.  That is the race check of the launch DAG (tests/test_gpu_train.py); it costs a
# couple of microseconds per reduction and serialises the weight-gradient GEMMs, so the default is the atomic form, whose
# results differ from run to run by the order of fp32 additions only.
DETERMINISTIC = os.environ.get("HG_DETERMINISTIC", "0") == "1"


def _pad(v: int, m: int) -> int:
    return (v + m - 1) // m * m


# ================================================================================================ parameters
class ParamStore:
    """Flat fp32 master parameters P, gradients G and RMSprop second moments V; the model's parameters
    (and their .grad) become views."""

    def __init__(self, model: nn.Module, device):
        self.device = torch.device(device)
        named = list(model.named_parameters())
        self.slots: Dict[str, tuple] = {}
        off = 0
        for name, p in named:
            self.slots[name] = (off, p.numel(), tuple(p.shape))
            off += _pad(p.numel(), 4)
        self.count = _pad(off, 4)
        total = self.count + 256          # tail padding: kernels may read a padded bias vector past the last tensor
        self.P = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.G = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.V = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.params = dict(named)
        with torch.no_grad():
            for name, p in named:
                view = self.view(self.P, name)
                view.copy_(p.detach())
                p.data = view
                p.grad = self.view(self.G, name)

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        off, n, shape = self.slots[name]
        t = flat[off:off + n]
        if len(shape) == 4:
            co, ci, kh, kw = shape
            return t.view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return t.view(shape)

    def flat(self, flat: torch.Tensor, name: str, extra: int = 0) -> torch.Tensor:
        off, n, _ = self.slots[name]
        return flat[off:off + n + extra]

    def rebind_grads(self):
        """optimizer.zero_grad(set_to_none=True) drops .grad: point it back at the flat gradient buffer."""
        for name, p in self.params.items():
            if p.grad is None or p.grad.data_ptr() != self.G.data_ptr() + self.slots[name][0] * 4:
                p.grad = self.view(self.G, name)


class _T:
    """An activation in the plan: bf16 NHWC data and (during backward emission) its gradient buffer.
    `can_stat`: the tensor is written by a GEMM whose epilogue can add the per-channel sums a train-mode BatchNorm
    needs; the first BatchNorm that consumes it sets `sums` (its own statistics slot) and the producer's launch --
    which looks `sums` up when it runs -- fills it, so no separate statistics pass re-reads the tensor.
    `can_pool` / `pool_out`: likewise for the 2x2 max-pool of the tensor (hg_conv_desc.pool_out): pool() sets pool_out
    and emits no launch of its own."""
    __slots__ = ("data", "grad", "grad_owned", "can_stat", "sums", "can_pool", "pool_out")

    def __init__(self, data, can_stat: bool = False, can_pool: bool = False):
        self.data = data
        self.grad = None
        self.grad_owned = True
        self.can_stat = can_stat
        self.sums = None
        self.can_pool = can_pool
        self.pool_out = None


class _Arena:
    """Gradient buffers of the backward pass, recycled in emission order.  With the launch DAG a recycled buffer
    is an edge (write-after-read) between otherwise independent launches, so reuse is first-in-first-out and only
    once `depth` buffers of that shape are free: the edge then points several bottlenecks back."""

    def __init__(self, device, depth: int = 1):
        self.device = device
        self.free = defaultdict(list)
        self.total_bytes = 0
        self.depth = depth

    def get(self, shape, dtype=None):
        dtype = dtype or _ACT
        key = (tuple(shape), dtype)
        if len(self.free[key]) >= self.depth:
            return self.free[key].pop(0)
        t = torch.empty(shape, dtype=dtype, device=self.device)
        self.total_bytes += t.numel() * t.element_size()
        return t

    def put(self, t):
        self.free[(tuple(t.shape), t.dtype)].append(t)

    def get_halo(self, n, h, w, c):
        key = ("halo", n, h, w, c)
        if len(self.free[key]) >= self.depth:
            return self.free[key].pop(0)
        t = ops.halo_padded_buffer(n, h, w, c, self.device)
        self.total_bytes += t.numel() * 2
        return t

    def put_halo(self, t, n, h, w, c):
        self.free[("halo", n, h, w, c)].append(t)


# ================================================================================================ records
class _Conv:
    """One nn.Conv2d: master slices, gradient slices, bf16 GEMM weights and its weight-pack entries."""

    def __init__(self, eng: "TrainEngine", conv: nn.Conv2d, *, dgrad: bool = True, fwd_ld: Optional[int] = None,
                 fwd_into=None):
        st = eng.store
        wname, bname = eng.pname[id(conv.weight)], eng.pname[id(conv.bias)]
        co, ci, kh, kw = conv.weight.shape
        # mobile=True: conv2 is depthwise (src/models/modules.py:15-17) -- a CUDA-core stencil straight on the fp32
        # master weights ([c][3][3][1] in the flat store = the kernel's [c][9]); no bf16 GEMM copies
        self.depthwise = conv.groups > 1 and conv.groups == conv.in_channels == co and (kh, kw) == (3, 3)
        if conv.groups != 1 and not self.depthwise:
            raise HgError(f"grouped convolution (groups={conv.groups}) outside the depthwise / concat_conv cases")
        self.co, self.ci, self.taps = co, ci, kh * kw
        self.row_len = self.taps * ci
        self.w, self.gw = st.flat(st.P, wname), st.flat(st.G, wname)
        co_pad = _pad(co, 16)
        self.b = st.flat(st.P, bname, extra=co_pad - co)      # padded view (tail of P is padded)
        self.gb = st.flat(st.G, bname)
        dev = st.device
        if self.depthwise:
            self.wf = self.wd = None
            return
        if fwd_into is not None:
            self.wf, col0, ld = fwd_into
        else:
            ld = fwd_ld or self.row_len
            self.wf, col0 = torch.zeros((co_pad, ld), dtype=_ACT, device=dev), 0
        self.wd = None
        dgrad_ld = 0
        if dgrad:
            dgrad_ld = _pad(self.taps * co, 64)
            self.wd = torch.zeros((ci, dgrad_ld), dtype=_ACT, device=dev)
        eng.pack_entries.append(dict(src=self.w, dst_fwd=self.wf, dst_dgrad=self.wd, co=co, taps=self.taps, ci=ci,
                                     fwd_ld=ld, fwd_col0=col0, dgrad_ld=dgrad_ld))


class _Bn:
    def __init__(self, eng: "TrainEngine", bn: nn.BatchNorm2d):
        st = eng.store
        self.c = bn.num_features
        self.gamma, self.beta = st.flat(st.P, eng.pname[id(bn.weight)]), st.flat(st.P, eng.pname[id(bn.bias)])
        self.ggamma, self.gbeta = st.flat(st.G, eng.pname[id(bn.weight)]), st.flat(st.G, eng.pname[id(bn.bias)])
        self.rm, self.rv, self.nbt = bn.running_mean, bn.running_var, bn.num_batches_tracked
        self.momentum = BN_MOMENTUM if bn.momentum is None else bn.momentum
        self.eps = bn.eps
        self.slot = eng.alloc_bn_slot(self.c)     # index into the per-step scratch (bound in bind())

    def bind(self, eng):
        o = self.slot
        c = self.c
        self.sums = eng.stat[o:o + 2 * c]             # forward: sum | sumsq
        self.bsums = eng.stat[o + 2 * c:o + 4 * c]    # backward: s1 | s2
        self.saved = eng.saved[o:o + 4 * c]           # mean | invstd | scale | shift


class _Block:
    """One HGBottleneck (src/models/modules.py:6-47): filled in by TrainEngine._block()."""
    __slots__ = ("cin", "planes", "cout", "bn1", "bn2", "bn3", "c1", "c2", "c3", "ds", "wf3", "b3")


class _Concat:
    """skip_mode='concat' (src/models/modules.py:58-61,91-93): conv1x1(cat([up1, upsample(low3)]), groups=2), ONE conv
    shared by the four levels of a stack's Hourglass.  Output channels [0,p) read up1, [p,2p) read low3; a 1x1
    convolution commutes with nearest upsampling, so it runs as two zero-padded 2p->2p GEMMs on the existing kernels:
    T = [0;W_b] low3 + [0;b_b] at low resolution, out = [W_a;0] up1 + [b_a;0] + upsample(T) (upsample-add epilogue)."""

    def __init__(self, eng: "TrainEngine", conv: nn.Conv2d):
        st = eng.store
        dev = st.device
        wname, bname = eng.pname[id(conv.weight)], eng.pname[id(conv.bias)]
        co, cig = conv.weight.shape[0], conv.weight.shape[1]
        if conv.groups != 2 or conv.kernel_size != (1, 1) or co % 2:
            raise HgError("concat_conv: expected a 1x1 convolution with groups=2")
        self.co, self.cig, self.half = co, cig, co // 2
        h = self.half
        self.w, self.gw = st.flat(st.P, wname), st.flat(st.G, wname)          # [co][cig]
        self.b, self.gb = st.flat(st.P, bname), st.flat(st.G, bname)
        self.wfa = torch.zeros((co, cig), dtype=_ACT, device=dev)
        self.wfb = torch.zeros((co, cig), dtype=_ACT, device=dev)
        self.wda = torch.zeros((cig, co), dtype=_ACT, device=dev)
        self.wdb = torch.zeros((cig, co), dtype=_ACT, device=dev)
        self.ba = torch.zeros(co, dtype=torch.float32, device=dev)
        self.bb = torch.zeros(co, dtype=torch.float32, device=dev)
        eng.pack_entries.append(dict(src=self.w[:h * cig], dst_fwd=self.wfa, dst_dgrad=self.wda, co=h, taps=1, ci=cig,
                                     fwd_ld=cig, fwd_col0=0, dgrad_ld=co))
        eng.pack_entries.append(dict(src=self.w[h * cig:], dst_fwd=self.wfb[h:], dst_dgrad=self.wdb[:, h:], co=h, taps=1,
                                     ci=cig, fwd_ld=cig, fwd_col0=0, dgrad_ld=co))
        eng.pack_entries.append(dict(src=self.b[:h], dst_f32=self.ba, co=h, taps=1, ci=1))
        eng.pack_entries.append(dict(src=self.b[h:], dst_f32=self.bb[h:], co=h, taps=1, ci=1))


class _Remap:
    """x + fc_(y) + score_(score(y)) (src/models/hourglass.py:86-89) as ONE merged 256->256 GEMM:
    Wm = W_fc_ + W_score_ W_score, bm = b_fc_ + b_score_ + W_score_ b_score; the chain rule back to the
    three convolutions' parameters runs in parameter space (hg_small_gemm_f32)."""

    def __init__(self, eng, fc_: nn.Conv2d, score_: nn.Conv2d, score: "_Conv"):
        st = eng.store
        dev = st.device
        n = eng.pname
        self.ch, self.J = fc_.out_channels, score_.in_channels
        self.wf_, self.gwf_ = st.flat(st.P, n[id(fc_.weight)]), st.flat(st.G, n[id(fc_.weight)])
        self.bf_, self.gbf_ = st.flat(st.P, n[id(fc_.bias)]), st.flat(st.G, n[id(fc_.bias)])
        self.ws_, self.gws_ = st.flat(st.P, n[id(score_.weight)]), st.flat(st.G, n[id(score_.weight)])
        self.bs_, self.gbs_ = st.flat(st.P, n[id(score_.bias)]), st.flat(st.G, n[id(score_.bias)])
        self.score = score
        ch = self.ch
        self.wm = torch.zeros(ch * ch, dtype=torch.float32, device=dev)
        self.bm = torch.zeros(ch, dtype=torch.float32, device=dev)
        self.wf = torch.zeros((ch, ch), dtype=_ACT, device=dev)
        self.wd = torch.zeros((ch, ch), dtype=_ACT, device=dev)
        eng.pack_entries.append(dict(src=self.wm, dst_fwd=self.wf, dst_dgrad=self.wd, co=ch, taps=1, ci=ch, fwd_ld=ch,
                                     fwd_col0=0, dgrad_ld=ch))

    def merge(self, ones):
        """Launches that build Wm / bm from the current masters (run before the weight pack)."""
        ch, J, s = self.ch, self.J, self.score
        ops.small_gemm(self.wm, self.ws_, s.w, self.wf_, ch, ch, J, J, 1, ch, 1, ch, 1)
        ops.small_gemm(self.bm, self.ws_, s.b, self.bf_, ch, 1, J, J, 1, 1, 0, 1, 0)
        ops.small_gemm(self.bm, self.bs_, ones, None, ch, 1, 1, 1, 0, 0, 0, 1, 0, beta=1.0)

    def chain(self, ones):
        """dWm sits in G[fc_.weight] (= dW_fc_), dbm in G[fc_.bias]; spread them to score_ / score."""
        ch, J, s = self.ch, self.J, self.score
        # dW_score_[i][j] = sum_k dWm[i][k] W_score[j][k] + dbm[i] b_score[j]
        ops.small_gemm(self.gws_, self.gwf_, s.w, None, ch, J, ch, ch, 1, 1, ch, J, 1)
        ops.small_gemm(self.gws_, self.gbf_, s.b, None, ch, J, 1, 1, 0, 0, 1, J, 1, beta=1.0)
        # dW_score[j][k] += sum_i W_score_[i][j] dWm[i][k] ;  db_score[j] += sum_i W_score_[i][j] dbm[i]
        ops.small_gemm(s.gw, self.ws_, self.gwf_, None, J, ch, ch, 1, J, ch, 1, ch, 1, beta=1.0)
        ops.small_gemm(s.gb, self.ws_, self.gbf_, None, J, 1, ch, 1, J, 1, 0, 1, 0, beta=1.0)
        # db_score_ = dbm
        ops.small_gemm(self.gbs_, self.gbf_, ones, None, ch, 1, 1, 1, 0, 0, 0, 1, 0)


# ================================================================================================ launch accounting
def _nbytes(*ts):
    return sum(t.numel() * t.element_size() for t in ts if torch.is_tensor(t))


def _describe(name, a, k):
    """Algorithmic flops / bytes of one library call, from its arguments (bench.py's roofline accounting)."""
    flops, nbytes, tag = 0.0, 0, name
    if name == "conv_nhwc":
        x, wt = a[0], a[1]
        pixels = x.numel() // x.shape[-1]
        cout = k["cout"]
        flops = 2.0 * pixels * wt.shape[1] * cout
        out = k.get("out") if k.get("out") is not None else k.get("out_nchw_f32")
        nbytes = _nbytes(x, wt, k.get("x2"), k.get("residual"), k.get("up_low"), out)
        tag = f"conv1x1_k{wt.shape[1]}_n{cout}_{x.shape[1]}x{x.shape[2]}" + ("_res" if k.get("residual") is not None else "") \
            + ("_up" if k.get("up_low") is not None else "") + ("_x2" if k.get("x2") is not None else "")
    elif name == "conv3x3_halo":
        wt = a[1]
        pixels = k["n"] * k["h"] * k["w"]
        flops = 2.0 * pixels * 9 * k["cin"] * k["cout"]
        nbytes = pixels * (k["cin"] + k["cout"]) * 2 + _nbytes(wt)
        tag = f"conv3x3h_k{9 * k['cin']}_n{k['cout']}_{k['h']}x{k['w']}"
    elif name == "wgrad":
        dout, z, dw = a[0], a[1], a[2]
        taps = k.get("taps", 1)
        rows = dout.numel() // dout.shape[-1]
        co = k.get("co_valid") or dout.shape[-1]
        flops = 2.0 * rows * co * z.shape[-1] * taps
        nbytes = _nbytes(dout, z) + co * z.shape[-1] * taps * 4
        tag = f"wgrad_co{dout.shape[-1]}_ci{z.shape[-1]}_t{taps}_rows{rows}"
    elif name in ("colstats", "bn_train_fwd", "bn_bwd_reduce", "bn_bwd_apply", "maxpool2x2", "maxpool2x2_bwd", "sumpool2x2",
                  "add_inplace", "nchw_to_nhwc_bf16_pad", "stem_im2col", "dwconv3x3", "dwconv3x3_wgrad"):
        big = [t for t in list(a) + list(k.values()) if torch.is_tensor(t) and t.numel() > 4096]
        nbytes = _nbytes(*big)
        ref = big[0] if big else None
        tag = name + (f"_c{ref.shape[-1]}_{ref.numel() // ref.shape[-1]}" if ref is not None and ref.dim() > 1 else "")
    elif name == "rmsprop_step":
        nbytes = 5 * _nbytes(a[0])
    return dict(op=tag, flops=flops, bytes=nbytes, kind="conv" if flops else "bw")


class _Recorder:
    """Wraps hgb200.ops during a plan's first eager pass to label every launch closure (no effect afterwards)."""

    def __init__(self, real):
        self.real = real
        self.calls: List[dict] = []

    def __getattr__(self, name):
        fn = getattr(self.real, name)
        if not callable(fn):
            return fn

        def wrapped(*a, **k):
            d = _describe(name, a, k)
            d["acc"] = dag.accesses(name, a, k)
            self.calls.append(d)
            return fn(*a, **k)

        return wrapped


# ================================================================================================ plan
class TrainPlan:
    def __init__(self, eng: "TrainEngine", n: int, h: int, w: int):
        self.eng = eng
        self.n, self.h, self.w = n, h, w
        dev = eng.device
        S, J = eng.num_stacks, eng.num_classes
        hh, ww = h // 4, w // 4
        self.input = torch.zeros((n, 3, h, w), dtype=torch.float32, device=dev)
        self.target = torch.zeros((n, J, hh, ww), dtype=torch.float32, device=dev)
        self.target_weight = torch.ones((n, J), dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.outputs = [torch.empty((n, J, hh, ww), dtype=torch.float32, device=dev) for _ in range(S)]
        self.dheat = [torch.zeros((n, J, hh, ww), dtype=torch.float32, device=dev) for _ in range(S)]
        self.pre: List[Callable] = []
        self.fwd: List[Callable] = []
        self.bwd: List[Callable] = []
        self.post: List[Callable] = []
        self.meta: List[dict] = []          # one entry per closure of launches("step"), filled by the first eager pass
        self.records: List[list] = []       # same alignment: tensor regions each closure reads / writes (hgb200/dag.py)
        self.schedules: Dict[str, tuple] = {}
        self.nodes: List[dict] = []
        self.grad_scale = 1.0          # 1/world_size: the sum over ranks is the global-batch mean
        self.use_target_weight = True
        self.graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self._dp_fn = None
        self.fwd_bytes = 0
        self.bwd_arena_bytes = 0
        self.pending_backward = False       # autograd drop-in: a forward whose backward has not run yet
        self.generation = 0                 # ... and which forward the static buffers currently belong to
        self.deterministic = False
        self.keep: List[torch.Tensor] = []
        self.scr = lambda t: None

    # ---- launch lists
    def _loss_launches(self):
        def run():
            ops.zero_(self.loss)
            ops.jmse_loss_into(self.outputs, self.dheat, self.target, self.target_weight if self.use_target_weight else None,
                               self.loss, grad_scale=self.grad_scale)
        return [run]

    def launches(self, which: str) -> List[Callable]:
        if which == "fwd":
            return self.pre + self.fwd
        if which == "bwd":
            return self.bwd + self.post
        if which == "step":
            return self.pre + self.fwd + self._loss_launches() + self.bwd + self.post
        raise KeyError(which)

    def run(self, which: str, use_graph: bool = True):
        if use_graph:
            g = self.graphs.get(which)
            if g is None:
                g = self._capture(which)
            g.replay()
        else:
            for fn in self.launches(which):
                fn()

    def profile(self, iters: int = 1):
        """Eager replay of the whole step with a CUDA-event pair around every closure (on the launching stream).
        Returns per-closure milliseconds averaged over `iters`, aligned with self.meta."""
        dev = self.eng.device
        st = torch.cuda.current_stream(dev)
        fns = self.launches("step")
        total = [0.0] * len(fns)
        for _ in range(iters):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(fns) + 1)]
            evs[0].record(st)
            for i, fn in enumerate(fns):
                fn()
                evs[i + 1].record(st)
            torch.cuda.synchronize(dev)
            for i in range(len(fns)):
                total[i] += evs[i].elapsed_time(evs[i + 1])
        return [t / iters for t in total]

    @property
    def num_kernel_launches(self) -> int:
        return sum(m.get("launches", 1) for m in self.meta)

    def _slice(self, which: str):
        """Index range of launches(which) inside launches("step") (meta / records are aligned with the latter)."""
        n_pre_fwd = len(self.pre) + len(self.fwd)
        n_loss = len(self._loss_launches())
        total = n_pre_fwd + n_loss + len(self.bwd) + len(self.post)
        return {"fwd": (0, n_pre_fwd), "bwd": (n_pre_fwd + n_loss, total), "step": (0, total)}[which]

    def schedule(self, which: str, streams: Optional[int] = None):
        """-> (dag, stream of each launch, cross-stream waits of each launch) for launches(which)."""
        k = streams or STREAMS
        sched = self.schedules.get((which, k))
        if sched is None:
            lo, hi = self._slice(which)
            d = dag.build(self.records[lo:hi])
            cost = [4e-6 + max(m["flops"] / 6e14, m["bytes"] / 3e12) for m in self.meta[lo:hi]]
            stream_of, waits = dag.assign_streams(d, cost, k, leaf_streams=LEAF_STREAMS)
            sched = self.schedules[(which, k)] = (d, stream_of, waits)
        return sched

    def run_dp(self, all_reduce: Callable, buckets):
        """The step with the per-bucket gradient all-reduce inside the same CUDA graph (see OVERLAP_ALLREDUCE)."""
        g = self.graphs.get("step_dp")
        if g is None or self._dp_fn != all_reduce:           # (bound methods compare equal by object and function)
            g = self._capture_dp(all_reduce, buckets, "step_dp")
            self._dp_fn = all_reduce
        g.replay()

    def comm_schedule(self, buckets):
        """-> (stream index, waits) of the all-reduce nodes appended to launches("step"): every node goes to ONE extra
        stream (index STREAMS) in bucket order and waits for the launches that last wrote its slice of the gradient buffer
        (RAW edges of the launch DAG, built from the same byte-range records as everything else)."""
        G = self.eng.store.G
        lo_, hi_ = self._slice("step")
        n_step = hi_ - lo_
        comm_records = [[([], [dag._region(G[lo:hi])])] for lo, hi in buckets]
        d = dag.build(self.records[lo_:hi_] + comm_records)
        return STREAMS, [sorted(p for p in d.preds[n_step + j] if p < n_step) for j in range(len(buckets))]

    def _capture_dp(self, all_reduce: Callable, buckets, key):
        fns = self.launches("step")
        G = self.eng.store.G
        _, stream_of, waits = self.schedule("step")
        comm_stream, comm_waits = self.comm_schedule(buckets)
        comm_fns = [(lambda lo=lo, hi=hi: all_reduce(G[lo:hi])) for lo, hi in buckets]
        prio = [-1] * (STREAMS - LEAF_STREAMS) + [0] * LEAF_STREAMS if (LEAF_STREAMS > 0 and STREAMS - LEAF_STREAMS >= 1) \
            else [0] * STREAMS
        g = dag.capture(fns + comm_fns, list(stream_of) + [comm_stream] * len(buckets), list(waits) + comm_waits,
                        STREAMS + 1, self.eng.device, priorities=prio + [-1])
        self.graphs[key] = g
        return g

    def _capture(self, which: str):
        # The caller must have run the list eagerly once before (module loading, shared-memory opt-in, access
        # records) -- TrainEngine.plan_for() does that on a scratch copy of the BN statistics.
        fns = self.launches(which)
        if STREAMS > 1:
            _, stream_of, waits = self.schedule(which)
            prio = None
            if LEAF_STREAMS > 0 and STREAMS - LEAF_STREAMS >= 1:
                prio = [-1] * (STREAMS - LEAF_STREAMS) + [0] * LEAF_STREAMS
            g = dag.capture(fns, stream_of, waits, STREAMS, self.eng.device, priorities=prio)
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for fn in fns:
                    fn()
        self.graphs[which] = g
        return g


class TrainEngine:
    def __init__(self, model: nn.Module, device=None):
        device = torch.device(device or next(model.parameters()).device)
        ops.require_device(device)            # raises off-CUDA: there is no CPU fallback
        self.device = device
        self.model = model
        for b in model.buffers():
            if b.device != device:
                raise HgError("move the model to the device (model.to('cuda')) before training")
        self.store = ParamStore(model, device)
        self.pname = {id(p): n for n, p in model.named_parameters()}
        self.pack_entries: List[dict] = []
        self._bn_floats = 0
        self._bns: List[_Bn] = []
        self.num_stacks = model.num_stacks
        self.num_classes = model.score[0].out_channels
        self.depth = 4

        self.stem = _Conv(self, model.conv1, dgrad=False, fwd_ld=192)
        self.stem_bn = self._bn(model.bn1)
        self.layer1 = [self._block(m) for m in model.layer1]
        self.layer2 = [self._block(m) for m in model.layer2]
        self.layer3 = [self._block(m) for m in model.layer3]
        self.hg, self.res, self.fc, self.score, self.remap, self.concat = [], [], [], [], [], []
        for i in range(self.num_stacks):
            self.concat.append(_Concat(self, model.hg[i].concat_conv) if hasattr(model.hg[i], "concat_conv") else None)
            levels = []
            for d in range(self.depth):
                levels.append([[self._block(m) for m in chain] for chain in model.hg[i].hg[d]])
            self.hg.append(levels)
            self.res.append([self._block(m) for m in model.res[i]])
            self.fc.append((_Conv(self, model.fc[i][0]), self._bn(model.fc[i][1])))
            self.score.append(_Conv(self, model.score[i]))
            if i < self.num_stacks - 1:
                self.remap.append(_Remap(self, model.fc_[i], model.score_[i], self.score[i]))
        # per-step scratch: BN forward/backward sums (zeroed every step) and saved statistics
        self.stat = torch.zeros(self._bn_floats, dtype=torch.float32, device=device)
        self.saved = torch.zeros(self._bn_floats, dtype=torch.float32, device=device)
        for b in self._bns:
            b.bind(self)
        self.ones = torch.ones(4, dtype=torch.float32, device=device)
        self.pack_table = ops.make_pack_table(self.pack_entries, device)
        self.plans: Dict[tuple, TrainPlan] = {}
        self.steps = 0
        self._buckets = None
        self._comm_warm = False

    # ------------------------------------------------------------------ construction helpers
    def alloc_bn_slot(self, c: int) -> int:
        o = self._bn_floats
        self._bn_floats += 4 * c
        return o

    def _bn(self, m):
        b = _Bn(self, m)
        self._bns.append(b)
        return b

    def _block(self, m):
        blk = _Block()
        blk.cin, blk.planes, blk.cout = m.conv1.in_channels, m.conv1.out_channels, m.conv3.out_channels
        blk.bn1, blk.bn2, blk.bn3 = self._bn(m.bn1), self._bn(m.bn2), self._bn(m.bn3)
        blk.c1, blk.c2 = _Conv(self, m.conv1), _Conv(self, m.conv2)
        blk.ds = None
        if m.downsample is not None:
            dev = self.device
            k = blk.planes + blk.cin
            blk.wf3 = torch.zeros((blk.cout, k), dtype=_ACT, device=dev)
            blk.c3 = _Conv(self, m.conv3, fwd_into=(blk.wf3, 0, k))
            blk.ds = _Conv(self, m.downsample[0], fwd_into=(blk.wf3, blk.planes, k))
            blk.b3 = torch.zeros(blk.cout, dtype=torch.float32, device=dev)
            self.pack_entries.append(dict(src=blk.c3.b[:blk.cout], src2=blk.ds.b[:blk.cout], dst_f32=blk.b3, co=blk.cout,
                                          taps=1, ci=1))
        else:
            blk.c3 = _Conv(self, m.conv3)
            blk.wf3 = blk.c3.wf
            blk.b3 = blk.c3.b
        return blk

    # ------------------------------------------------------------------ plan construction
    def build_plan(self, n: int, h: int, w: int) -> TrainPlan:
        if h % 64 or w % 64:
            raise HgError(f"input {h}x{w}: height and width must be multiples of 64")
        plan = TrainPlan(self, n, h, w)
        dev = self.device
        F, nodes = plan.fwd, plan.nodes
        fwd_bytes = [0]
        plan.deterministic = DETERMINISTIC

        def scr(t):
            """Per-launch-site scratch of the fixed-order reductions over the pixels of NHWC tensor `t` (None: atomics)."""
            if not plan.deterministic:
                return None
            s = ops.colreduce_scratch(t.numel() // t.shape[-1], t.shape[-1], dev)
            plan.keep.append(s)
            return s

        plan.scr = scr

        def new(shape, dtype=None):
            dtype = dtype or _ACT
            t = torch.empty(shape, dtype=dtype, device=dev)
            fwd_bytes[0] += t.numel() * t.element_size()
            return t

        def new_halo(nn_, hh_, ww_, c):
            t = ops.halo_padded_buffer(nn_, hh_, ww_, c, dev)
            fwd_bytes[0] += t.numel() * 2
            return t

        def bn_fwd(bn: _Bn, x, out, halo=False, have_stats=False):
            """relu(bn(x)) with batch statistics.  x: raw tensor (have_stats: its producer was handed bn.sums) or a _T
            (whose producer fills the sums when it can)."""
            if isinstance(x, _T):
                if x.can_stat and x.sums is None and x.pool_out is None:
                    x.sums, have_stats = bn.sums, True
                x = x.data
            if not have_stats:
                # sums about x[pixel 0]: no cancellation in E[x^2] - E[x]^2 when |mean| >> std
                s = scr(x)
                F.append(lambda: ops.colstats(x, bn.sums[:bn.c], bn.sums[bn.c:], shift=True, scratch=s))
            shifted = not have_stats
            F.append(lambda: ops.bn_train_fwd(x, bn.sums, bn.gamma, bn.beta, bn.rm, bn.rv, bn.nbt, bn.saved, out, halo=halo,
                                              relu=True, eps=bn.eps, momentum=bn.momentum, shifted=shifted))

        def block(blk: _Block, x: _T, up_low: Optional[_T] = None) -> _T:
            nb, hh_, ww_, cin = x.data.shape
            pl = blk.planes
            z1 = new((nb, hh_, ww_, cin))
            a1 = new((nb, hh_, ww_, pl))
            z2h = new((nb, hh_, ww_, pl)) if blk.c2.depthwise else new_halo(nb, hh_, ww_, pl)
            a2 = new((nb, hh_, ww_, pl))
            z3 = new((nb, hh_, ww_, pl))
            fs = FUSED_STATS and nb * hh_ * ww_ <= FUSED_STATS_MAX_PIXELS
            y = _T(new((nb, hh_, ww_, blk.cout)), can_stat=fs,
                   can_pool=FUSE_POOL and not fs and ops.conv_pool_fusable(hh_, ww_, blk.cout, pl + (cin if blk.ds is not None else 0)))
            xd = x.data
            bn_fwd(blk.bn1, x, z1)
            F.append(lambda: ops.conv_nhwc(z1, blk.c1.wf, blk.c1.b, ksize=1, cout=pl, out=a1,
                                           stats=blk.bn2.sums if fs else None))
            if blk.c2.depthwise:
                bn_fwd(blk.bn2, a1, z2h, have_stats=fs)
                F.append(lambda: ops.dwconv3x3(z2h, blk.c2.w, blk.c2.b, out=a2))
                bn_fwd(blk.bn3, a2, z3)
            else:
                bn_fwd(blk.bn2, a1, z2h, halo=True, have_stats=fs)
                F.append(lambda: ops.conv3x3_halo(z2h, blk.c2.wf, blk.c2.b, n=nb, h=hh_, w=ww_, cin=pl, cout=pl, out=a2,
                                                  stats=blk.bn3.sums if fs else None))
                bn_fwd(blk.bn3, a2, z3, have_stats=fs)
            lowd = up_low.data if up_low is not None else None
            if blk.ds is not None:
                F.append(lambda: ops.conv_nhwc(z3, blk.wf3, blk.b3, ksize=1, cout=blk.cout, x2=xd, up_low=lowd, out=y.data,
                                               stats=y.sums, pool_out=y.pool_out))
            else:
                F.append(lambda: ops.conv_nhwc(z3, blk.wf3, blk.b3, ksize=1, cout=blk.cout, residual=xd, up_low=lowd,
                                               out=y.data, stats=y.sums, pool_out=y.pool_out))
            nodes.append(dict(kind="block", blk=blk, x=x, y=y, up_low=up_low, z1=z1, a1=a1, z2h=z2h, a2=a2, z3=z3))
            return y

        def chain(blocks, x, up_low=None):
            cur = x
            for i, blk in enumerate(blocks):
                cur = block(blk, cur, up_low if i == len(blocks) - 1 else None)
            return cur

        def pool(x: _T) -> _T:
            nb, hh_, ww_, c = x.data.shape
            p = _T(new((nb, hh_ // 2, ww_ // 2, c)))
            if x.can_pool and x.pool_out is None and x.sums is None:
                x.pool_out = p.data          # x's producer (looked up when it runs) writes the pooled tensor itself
                x.can_stat = False           # the kernel does one or the other
            else:
                F.append(lambda: ops.maxpool2x2(x.data, p.data))
            nodes.append(dict(kind="pool", x=x, y=p))
            return p

        def hourglass(levels, d, x: _T, cat: Optional[_Concat]) -> _T:
            p = pool(x)
            low1 = chain(levels[d][1], p)
            low2 = hourglass(levels, d - 1, low1, cat) if d > 0 else chain(levels[0][3], low1)
            low3 = chain(levels[d][2], low2)
            if cat is None:
                return chain(levels[d][0], x, up_low=low3)
            up1 = chain(levels[d][0], x)
            t = new(low3.data.shape)
            fsc = FUSED_STATS and x.data.numel() // x.data.shape[-1] <= FUSED_STATS_MAX_PIXELS
            y = _T(new(x.data.shape), can_stat=fsc,
                   can_pool=FUSE_POOL and not fsc and ops.conv_pool_fusable(x.data.shape[1], x.data.shape[2], cat.co, cat.cig))
            F.append(lambda: ops.conv_nhwc(low3.data, cat.wfb, cat.bb, ksize=1, cout=cat.co, out=t))
            F.append(lambda: ops.conv_nhwc(up1.data, cat.wfa, cat.ba, ksize=1, cout=cat.co, up_low=t, out=y.data,
                                           stats=y.sums, pool_out=y.pool_out))
            nodes.append(dict(kind="concat", cat=cat, up1=up1, low3=low3, y=y))
            return y

        # ---- stem: im2col rows (kept: they are the stem's wgrad operand) -> GEMM -> BN+ReLU
        rows = new((n, h // 2, w // 2, 192))
        a0 = new((n, h // 2, w // 2, 64))
        s0 = _T(new((n, h // 2, w // 2, 64)))
        F.append(lambda: ops.stem_im2col(plan.input, out=rows))
        fs0 = FUSED_STATS and n * (h // 2) * (w // 2) <= FUSED_STATS_MAX_PIXELS
        F.append(lambda: ops.conv_nhwc(rows, self.stem.wf, self.stem.b, ksize=1, cout=64, out=a0,
                                       stats=self.stem_bn.sums if fs0 else None))
        bn_fwd(self.stem_bn, a0, s0.data, have_stats=fs0)
        nodes.append(dict(kind="stem", rows=rows, a0=a0, y=s0))
        l1 = chain(self.layer1, s0)
        p1 = pool(l1)
        l2 = chain(self.layer2, p1)
        x = chain(self.layer3, l2)
        J = self.num_classes
        for i in range(self.num_stacks):
            y = hourglass(self.hg[i], self.depth - 1, x, self.concat[i])
            y = chain(self.res[i], y)
            fcc, fcb = self.fc[i]
            nb, hh_, ww_, ch = y.data.shape
            afc = new((nb, hh_, ww_, ch))
            y2 = _T(new((nb, hh_, ww_, ch)))
            fsh = FUSED_STATS and nb * hh_ * ww_ <= FUSED_STATS_MAX_PIXELS
            F.append(lambda y=y, afc=afc, fcc=fcc, fcb=fcb, fsh=fsh: ops.conv_nhwc(y.data, fcc.wf, fcc.b, ksize=1, cout=ch,
                                                                                   out=afc, stats=fcb.sums if fsh else None))
            bn_fwd(fcb, afc, y2.data, have_stats=fsh)
            nodes.append(dict(kind="fc", conv=fcc, bn=fcb, x=y, a=afc, y=y2))
            sc = self.score[i]
            F.append(lambda y2=y2, sc=sc, i=i: ops.conv_nhwc(y2.data, sc.wf, sc.b, ksize=1, cout=J, heads=True,
                                                             out_nchw_f32=plan.outputs[i]))
            nodes.append(dict(kind="score", conv=sc, x=y2, idx=i))
            if i < self.num_stacks - 1:
                rm = self.remap[i]
                xn = _T(new((nb, hh_, ww_, ch)), can_stat=fsh,
                        can_pool=FUSE_POOL and not fsh and ops.conv_pool_fusable(hh_, ww_, ch, ch))
                F.append(lambda y2=y2, rm=rm, x=x, xn=xn: ops.conv_nhwc(y2.data, rm.wf, rm.bm, ksize=1, cout=ch,
                                                                       residual=x.data, out=xn.data, stats=xn.sums,
                                                                       pool_out=xn.pool_out))
                nodes.append(dict(kind="remap", rm=rm, x=x, y2=y2, y=xn))
                x = xn
        plan.fwd_bytes = fwd_bytes[0]

        # ---- pre: zero the per-step scratch, merge the remap weights, pack every GEMM weight (one launch)
        plan.pre.append(lambda: ops.zero_(self.stat))
        for rm in self.remap:
            plan.pre.append(lambda rm=rm: rm.merge(self.ones))
        # the pack launch declares what it touches, so that launches that need no weights (input layout change, memsets)
        # are not ordered after it in the launch DAG
        # ... and it is split in two: the forward layouts (coalesced, ~50 us), which the forward pass waits for, and the
        # transposed dgrad layouts (scattered 2-byte writes, ~250 us), which only the backward pass needs
        reads = [e[k] for e in self.pack_entries for k in ("src", "src2") if e.get(k) is not None]
        w_fwd = [e[k] for e in self.pack_entries for k in ("dst_f32", "dst_fwd") if e.get(k) is not None]
        w_dg = [e["dst_dgrad"] for e in self.pack_entries if e.get("dst_dgrad") is not None]
        plan.pre.append(lambda: ops.pack_weights(self.pack_table, len(self.pack_entries), reads=reads, writes=w_fwd, which=1))
        plan.pre.append(lambda: ops.pack_weights(self.pack_table, len(self.pack_entries), reads=reads, writes=w_dg, which=2))

        self._emit_backward(plan)
        return plan

    # ------------------------------------------------------------------ backward emission (reverse node order)
    def _emit_backward(self, plan: TrainPlan):
        dev = self.device
        B = plan.bwd
        arena = _Arena(dev, depth=3 if STREAMS > 1 else 1)
        G = self.store.G
        B.append(lambda: ops.zero_(G))

        def release(t: _T):
            if t.grad is not None and t.grad_owned:
                arena.put(t.grad)
            t.grad = None

        def dgrad1x1(g, wd, cout, out, residual=None):
            B.append(lambda: ops.conv_nhwc(g, wd, None, ksize=1, cout=cout, residual=residual, out=out))

        scr = plan.scr
        wg_ctas = 1 if plan.deterministic else 0

        def bn_bwd(bn: _Bn, dz, xdata, out, *, add1=None, add2=None, halo=False):
            s = scr(xdata)
            B.append(lambda: ops.bn_bwd_reduce(dz, xdata, bn.saved, bn.bsums, scratch=s))
            B.append(lambda: ops.bn_bwd_apply(dz, xdata, bn.saved, bn.bsums, out, add1=add1, add2=add2, dgamma=bn.ggamma,
                                              dbeta=bn.gbeta, halo=halo))

        for node in reversed(plan.nodes):
            kind = node["kind"]
            if kind == "block":
                blk, x, y, low = node["blk"], node["x"], node["y"], node["up_low"]
                gy = y.grad
                if gy is None:
                    raise HgError("internal: block output without a gradient")
                nb, hh_, ww_, cin = x.data.shape
                pl, cout = blk.planes, blk.cout
                if low is not None:
                    acc = low.grad is not None
                    if not acc:
                        low.grad = arena.get(low.data.shape)
                    B.append(lambda gy=gy, lg=low.grad, acc=acc: ops.sumpool2x2(gy, lg, accumulate=acc))
                B.append(lambda gy=gy, blk=blk, s=scr(gy): ops.colstats(gy, blk.c3.gb, scratch=s))
                B.append(lambda gy=gy, blk=blk, z3=node["z3"]: ops.wgrad(gy, z3, blk.c3.gw, max_ctas=wg_ctas))
                if blk.ds is not None:
                    B.append(lambda gy=gy, blk=blk, s=scr(gy): ops.colstats(gy, blk.ds.gb, scratch=s))
                    B.append(lambda gy=gy, blk=blk, xd=x.data: ops.wgrad(gy, xd, blk.ds.gw, max_ctas=wg_ctas))
                dz3 = arena.get((nb, hh_, ww_, pl))
                dgrad1x1(gy, blk.c3.wd, pl, dz3)
                if blk.c2.depthwise:
                    da2 = arena.get((nb, hh_, ww_, pl))
                    bn_bwd(blk.bn3, dz3, node["a2"], da2)
                    arena.put(dz3)
                    B.append(lambda da2=da2, z2=node["z2h"], blk=blk: ops.dwconv3x3_wgrad(da2, z2, blk.c2.gw))
                    dz2 = arena.get((nb, hh_, ww_, pl))
                    B.append(lambda da2=da2, blk=blk, dz2=dz2: ops.dwconv3x3(da2, blk.c2.w, None, flip=True, out=dz2))
                    arena.put(da2)
                else:
                    da2h = arena.get_halo(nb, hh_, ww_, pl)
                    bn_bwd(blk.bn3, dz3, node["a2"], da2h, halo=True)
                    arena.put(dz3)
                    B.append(lambda da2h=da2h, z2h=node["z2h"], blk=blk, P=ww_ + 1, pl=pl:
                             ops.wgrad(da2h.view(-1, pl), z2h.view(-1, pl), blk.c2.gw, taps=9, halo_pitch=P, max_ctas=wg_ctas))
                    dz2 = arena.get((nb, hh_, ww_, pl))
                    B.append(lambda da2h=da2h, blk=blk, dz2=dz2, nb=nb, hh_=hh_, ww_=ww_, pl=pl:
                             ops.conv3x3_halo(da2h, blk.c2.wd, None, n=nb, h=hh_, w=ww_, cin=pl, cout=pl, out=dz2))
                    arena.put_halo(da2h, nb, hh_, ww_, pl)
                da1 = arena.get((nb, hh_, ww_, pl))
                bn_bwd(blk.bn2, dz2, node["a1"], da1)
                arena.put(dz2)
                B.append(lambda da1=da1, z1=node["z1"], blk=blk: ops.wgrad(da1, z1, blk.c1.gw, max_ctas=wg_ctas))
                dz1 = arena.get((nb, hh_, ww_, cin))
                dgrad1x1(da1, blk.c1.wd, cin, dz1)
                arena.put(da1)
                # gradient of the block input: BN1 backward + the residual path (+ whatever x already collected)
                if blk.ds is not None:
                    r = arena.get((nb, hh_, ww_, cin))
                    dgrad1x1(gy, blk.ds.wd, cin, r)
                    if x.grad is None:
                        x.grad = arena.get((nb, hh_, ww_, cin))
                        bn_bwd(blk.bn1, dz1, x.data, x.grad, add1=r)
                    else:
                        bn_bwd(blk.bn1, dz1, x.data, x.grad, add1=r, add2=x.grad)
                    arena.put(r)
                    release(y)
                else:
                    if x.grad is None:
                        # the identity path hands y's gradient buffer on to x; BN1's contribution is added in place
                        x.grad, x.grad_owned = gy, y.grad_owned
                        y.grad = None
                        bn_bwd(blk.bn1, dz1, x.data, x.grad, add1=gy)
                    else:
                        bn_bwd(blk.bn1, dz1, x.data, x.grad, add1=gy, add2=x.grad)
                        release(y)
                arena.put(dz1)
            elif kind == "concat":
                cat, up1, low3, y = node["cat"], node["up1"], node["low3"], node["y"]
                gy = y.grad
                h_ = cat.half
                B.append(lambda gy=gy, cat=cat, s=scr(gy): ops.colstats(gy, cat.gb, scratch=s))
                B.append(lambda gy=gy, cat=cat, u=up1.data, h_=h_: ops.wgrad(gy, u, cat.gw, co_valid=h_, max_ctas=wg_ctas))
                dt = arena.get(low3.data.shape)
                B.append(lambda gy=gy, dt=dt: ops.sumpool2x2(gy, dt, accumulate=False))
                B.append(lambda dt=dt, cat=cat, l3=low3.data, h_=h_:
                         ops.wgrad(dt, l3, cat.gw[h_ * cat.cig:], co_first=h_, max_ctas=wg_ctas))
                assert up1.grad is None and low3.grad is None
                up1.grad = arena.get(up1.data.shape)
                dgrad1x1(gy, cat.wda, cat.cig, up1.grad)
                low3.grad = arena.get(low3.data.shape)
                dgrad1x1(dt, cat.wdb, cat.cig, low3.grad)
                arena.put(dt)
                release(y)
            elif kind == "pool":
                x, y = node["x"], node["y"]
                acc = x.grad is not None
                if not acc:
                    x.grad = arena.get(x.data.shape)
                B.append(lambda x=x.data, gy=y.grad, gx=x.grad, acc=acc: ops.maxpool2x2_bwd(x, gy, gx, acc))
                release(y)
            elif kind == "remap":
                rm, x, y2, y = node["rm"], node["x"], node["y2"], node["y"]
                gy = y.grad
                B.append(lambda gy=gy, rm=rm, s=scr(gy): ops.colstats(gy, rm.gbf_, scratch=s))
                B.append(lambda gy=gy, rm=rm, y2=y2.data: ops.wgrad(gy, y2, rm.gwf_, max_ctas=wg_ctas))
                assert y2.grad is None
                y2.grad = arena.get(y2.data.shape)
                dgrad1x1(gy, rm.wd, rm.ch, y2.grad)
                assert x.grad is None
                x.grad, x.grad_owned = gy, y.grad_owned      # identity: x inherits the buffer
                y.grad = None
            elif kind == "score":
                sc, x, i = node["conv"], node["x"], node["idx"]
                nb, hh_, ww_, ch = x.data.shape
                J = self.num_classes
                dhp = arena.get((nb, hh_, ww_, 64))
                B.append(lambda i=i, dhp=dhp: ops.nchw_to_nhwc_bf16_pad(plan.dheat[i], dhp))
                B.append(lambda dhp=dhp, sc=sc, J=J, s=scr(dhp): ops.colstats(dhp, sc.gb, c_valid=J, scratch=s))
                B.append(lambda dhp=dhp, sc=sc, xd=x.data, J=J: ops.wgrad(dhp, xd, sc.gw, co_valid=J, max_ctas=wg_ctas))
                # the merged remap conv that follows this head has already left dWm / dbm (its backward runs first): spread
                # them to score_ / score HERE, not at the end of the step, so that this stack's slice of the gradient
                # buffer is final -- and can be exchanged -- as soon as its backward is
                for rm in self.remap:
                    if rm.score is sc:
                        B.append(lambda rm=rm: rm.chain(self.ones))
                if x.grad is None:
                    x.grad = arena.get(x.data.shape)
                    dgrad1x1(dhp, sc.wd, ch, x.grad)
                else:
                    dgrad1x1(dhp, sc.wd, ch, x.grad, residual=x.grad)
                arena.put(dhp)
            elif kind == "fc":
                cv, bn, x, y = node["conv"], node["bn"], node["x"], node["y"]
                da = arena.get(y.data.shape)
                bn_bwd(bn, y.grad, node["a"], da)
                release(y)
                B.append(lambda da=da, xd=x.data, cv=cv: ops.wgrad(da, xd, cv.gw, max_ctas=wg_ctas))
                assert x.grad is None
                x.grad = arena.get(x.data.shape)
                dgrad1x1(da, cv.wd, cv.ci, x.grad)
                arena.put(da)
            elif kind == "stem":
                y = node["y"]
                da0 = arena.get(y.data.shape)
                bn_bwd(self.stem_bn, y.grad, node["a0"], da0)
                release(y)
                B.append(lambda da0=da0, rows=node["rows"]: ops.wgrad(da0, rows, self.stem.gw, ci_valid=147, ld=147,
                                                                      tap_stride=147, max_ctas=wg_ctas))
                arena.put(da0)
            else:
                raise HgError(f"internal: unknown node kind {kind}")
        plan.bwd_arena_bytes = arena.total_bytes

    # ------------------------------------------------------------------ public
    def plan_for(self, n: int, h: int, w: int) -> TrainPlan:
        key = (n, h, w)
        p = self.plans.get(key)
        if p is None:
            p = self.build_plan(n, h, w)
            # one eager pass with the BN running statistics preserved: loads every kernel and opts in to large
            # shared memory before any graph capture, and records what every launch reads and writes (launch DAG,
            # roofline accounting).  On a side stream when on the GPU.
            cuda = self.device.type == "cuda"
            bufs = [(b.rm, b.rv, b.nbt) for b in self._bns]
            keep = [(a.clone(), b.clone(), c.clone()) for a, b, c in bufs]
            global ops
            guard = dag.RECORDING           # the rebinding of `ops` below is process-wide: one recording at a time
            guard.__enter__()
            rec = _Recorder(ops)
            real_ops, ops = ops, rec

            def record():
                for fn in p.launches("step"):
                    rec.calls = []
                    fn()
                    c = rec.calls
                    p.meta.append(dict(op=c[0]["op"] if c else "memset", flops=sum(x["flops"] for x in c),
                                       bytes=sum(x["bytes"] for x in c), kind=c[0]["kind"] if c else "bw",
                                       launches=len(c)))
                    p.records.append([x["acc"] for x in c])
            try:
                if cuda:
                    s = torch.cuda.Stream(device=self.device)
                    s.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(s):
                        record()
                    torch.cuda.current_stream(self.device).wait_stream(s)
                    torch.cuda.synchronize(self.device)
                else:
                    record()
            finally:
                ops = real_ops
                guard.__exit__(None, None, None)
            if cuda:
                ops.check_err_word(self.device)
            for (a, b, c), (ka, kb, kc) in zip(bufs, keep):
                a.copy_(ka)
                b.copy_(kb)
                c.copy_(kc)
            self.plans[key] = p
        return p

    def rmsprop(self, lr: float, alpha: float = RMSPROP_ALPHA, eps: float = RMSPROP_EPS):
        st = self.store
        ops.rmsprop_step(st.P[:st.count], st.G[:st.count], st.V[:st.count], lr, alpha, eps)
        self.mark_updated()

    def mark_updated(self):
        self.steps += 1
        self.model._weights_epoch = getattr(self.model, "_weights_epoch", 0) + 1

    def release_graphs(self):
        """Drop every captured CUDA graph (they are re-captured on demand).  Required before
        torch.distributed.destroy_process_group() when the step's graph holds the all-reduce nodes."""
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        for p in self.plans.values():
            p.graphs.clear()
            p._dp_fn = None

    def grad_buckets(self):
        """Slices [lo, hi) of the flat gradient buffer in the order the backward pass completes them: the hourglass of the
        last stack first (13 of a stack's 14 bottlenecks live in `hg.<i>`, contiguous in named_parameters() order), then
        down to stack 1, then everything else (stem, layer1-3, res / fc / score / remap convs: ~12 % of the parameters).
        Identical on every rank (it depends on the parameter names only), so the ranks issue the same collectives in
        the same order whether they run the graph, the eager list or idle_step()."""
        if self._buckets is None:
            st = self.store
            spans = {}
            for name, (off, n, _) in st.slots.items():
                parts = name.split(".")
                if parts[0] == "hg" and len(parts) > 2 and parts[1].isdigit():
                    lo, hi = spans.get(int(parts[1]), (off, off))
                    spans[int(parts[1])] = (min(lo, off), max(hi, off + _pad(n, 4)))
            out, covered = [], []
            for i in sorted(spans, reverse=True):
                out.append(spans[i])
                covered.append(spans[i])
            covered.sort()
            # the hourglass spans must be disjoint runs; whatever lies between / around them forms the closing buckets
            pos, rest = 0, []
            for lo, hi in covered:
                if lo < pos:
                    out, rest, pos = [], [], 0           # unexpected layout: one bucket
                    break
                if lo > pos:
                    rest.append((pos, lo))
                pos = hi
            if pos < st.count:
                rest.append((pos, st.count))
            # order of completion in the backward pass: hg.<S-1> ... hg.<1>, then the block BEHIND the hourglasses (res / fc /
            # score / remap convs of every stack: stack 1's are written before ITS hourglass is entered), then hg.<0>, and
            # last the block in FRONT of them (stem, layer1-3) -- the only bucket whose exchange nothing can hide
            if out and len(rest) == 2 and rest[0][0] == 0:
                self._buckets = out[:-1] + [rest[1], out[-1], rest[0]]
            else:
                self._buckets = out + rest
        return self._buckets

    def _reduce_eager(self, all_reduce: Callable):
        G = self.store.G
        if OVERLAP_ALLREDUCE:
            for lo, hi in self.grad_buckets():
                all_reduce(G[lo:hi])
        else:
            all_reduce(G[:self.store.count])

    def _comm_warmup(self, all_reduce: Callable):
        """The communicator is created by the first collective: that must not happen under stream capture.  Every rank
        passes here exactly once (its first train_step or idle_step), so the extra collective is symmetric."""
        if OVERLAP_ALLREDUCE and not self._comm_warm:
            self._comm_warm = True
            all_reduce(torch.zeros(4, dtype=torch.float32, device=self.device))

    def idle_step(self, lr: float, all_reduce: Optional[Callable] = None):
        """A rank whose shard of a (ragged, last) batch is empty still takes part in the step's collectives: zero
        gradients in, the other ranks' sum out, the same RMSprop update as everywhere else."""
        ops.zero_(self.store.G)
        if all_reduce is not None:
            self._comm_warmup(all_reduce)
            self._reduce_eager(all_reduce)
        self.rmsprop(lr)

    def train_step(self, x: torch.Tensor, target: torch.Tensor, target_weight: Optional[torch.Tensor], lr: float, *,
                   use_graph: bool = True, world_size: int = 1, all_reduce: Optional[Callable] = None,
                   grad_scale: Optional[float] = None) -> torch.Tensor:
        """One fused step: forward, JointsMSE loss, backward, (all-reduce,) RMSprop.  Returns the plan's loss
        tensor (device, fp32 [1]): the mean loss of THIS rank's shard; only the gradients carry the 1/world_size
        factor (or `grad_scale` = shard size / global batch size when the shards are uneven), so that their sum over
        ranks is the gradient of the global-batch mean."""
        n, _, h, w = x.shape
        plan = self.plan_for(n, h, w)
        plan.input.copy_(x, non_blocking=True)
        plan.target.copy_(target, non_blocking=True)
        if target_weight is not None:
            plan.target_weight.copy_(target_weight.reshape(n, -1), non_blocking=True)
        gs = 1.0 / world_size if grad_scale is None else float(grad_scale)
        if plan.grad_scale != gs or plan.use_target_weight != (target_weight is not None):
            plan.grad_scale, plan.use_target_weight = gs, target_weight is not None
            plan.graphs.clear()
        if all_reduce is not None:
            self._comm_warmup(all_reduce)
        if all_reduce is not None and OVERLAP_ALLREDUCE and use_graph and STREAMS > 1:
            plan.run_dp(all_reduce, self.grad_buckets())
        else:
            plan.run("step", use_graph)
            if all_reduce is not None:
                self._reduce_eager(all_reduce)
        self.rmsprop(lr)
        return plan.loss


# ================================================================================================ autograd drop-in
class _TrainFn(torch.autograd.Function):
    """model(x) in train mode under the reference's own loop (criterion(...).backward(); optimizer.step()):
    forward runs the plan's forward graph; backward feeds the heat-map gradients to the backward graph, which
    leaves every parameter gradient in the flat buffer the parameters' .grad views alias."""

    @staticmethod
    def forward(ctx, x, anchor, eng: TrainEngine, use_graph: bool):
        n, _, h, w = x.shape
        plan = eng.plan_for(n, h, w)
        # The saved activations live in the plan's static buffers: a second forward of the same shape overwrites what the
        # backward of the first would read.  A forward whose backward never runs (logging, evaluation in train mode) is
        # legitimate, so re-entry is allowed and every forward takes a new generation number; backward() refuses to run on
        # a forward that is no longer the plan's latest (gradient accumulation over micro-batches of ONE shape is not
        # supported on this path: use a larger batch, or call backward before the next forward).
        plan.generation += 1
        plan.input.copy_(x, non_blocking=True)
        plan.run("fwd", use_graph)
        plan.pending_backward = True
        ctx.plan, ctx.eng, ctx.use_graph, ctx.generation = plan, eng, use_graph, plan.generation
        return tuple(o.clone() for o in plan.outputs)

    @staticmethod
    def backward(ctx, *grads):
        plan, eng = ctx.plan, ctx.eng
        if ctx.generation != plan.generation:
            raise HgError("HourglassNet backward: another forward of the same shape has run since this one; its saved "
                          "activations (the plan's static buffers) are gone")
        for dst, g in zip(plan.dheat, grads):
            if g is None:
                dst.zero_()
            else:
                dst.copy_(g, non_blocking=True)
        plan.run("bwd", ctx.use_graph)
        plan.pending_backward = False
        eng.store.rebind_grads()
        return None, None, None, None


def train_engine(model: nn.Module) -> TrainEngine:
    eng = getattr(model, "_train_engine", None)
    if eng is None:
        eng = TrainEngine(model)
        model._train_engine = eng
    return eng


def training_forward(model: nn.Module, x: torch.Tensor):
    dev = next(model.parameters()).device
    try:
        ops.require_device(dev)
    except HgError as e:
        raise RuntimeError(f"HourglassNet (B200 build) trains on CUDA only: {e}") from None
    eng = train_engine(model)
    x = x.to(device=dev, dtype=torch.float32).contiguous()
    if not torch.is_grad_enabled():
        n, _, h, w = x.shape
        plan = eng.plan_for(n, h, w)
        plan.input.copy_(x, non_blocking=True)
        plan.run("fwd", model.use_cuda_graph)
        return [o.clone() for o in plan.outputs]
    anchor = model.conv1.weight          # any leaf that requires grad: makes autograd call backward()
    model._weights_epoch = getattr(model, "_weights_epoch", 0) + 1     # an optimizer step will follow
    return list(_TrainFn.apply(x, anchor, eng, model.use_cuda_graph))
