"""Launch DAG of a static plan: which launch must wait for which, and how the list is captured across
several CUDA streams so that one CUDA graph carries the true dependencies instead of one long chain.

Why: at 32 images per GPU a training step is ~2800 launches, three quarters of them on the 32x32 .. 4x4 levels
of the hourglass where a kernel cannot fill 148 SMs.  The weight-gradient GEMMs and bias-gradient sums are
leaves of the backward pass, the `up1` bottleneck of every hourglass level is independent of the whole lower
pyramid (src/models/modules.py:80-96), and so on: captured as a DAG those run side by side.

  * `accesses(name, args, kwargs)`: the tensors one library call reads and writes (table below).
  * `build(records)`: RAW / WAR / WAW edges from byte-range overlap of those tensors, transitively reduced.
  * `assign_streams(...)`: list scheduling onto K streams -- a launch follows its most recent predecessor on
    that predecessor's stream whenever it is still the stream's tail (keeps programmatic dependent launch along
    chains), otherwise takes a stream whose tail is already one of its ancestors (no false edge), otherwise the
    stream that frees up first.
  * `capture(...)`: replays the closures under `torch.cuda.graph`, switching streams and wiring events.

The analysis is host logic only (tests/test_train_dag_cpu.py executes random topological orders of the DAG
through the CPU emulation and requires bit-identical results)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# positional / keyword arguments a call WRITES (everything else that is a tensor is read).  "rw" entries are
# read-modify-write (accumulations): they order like writes.
_WRITES: Dict[str, Tuple[Tuple, Tuple]] = {
    # name: (written positional indices, written keyword names)
    "conv_nhwc": ((), ("out", "out_nchw_f32", "out_halo", "stats", "pool_out", "pool_in")),
    "conv3x3_halo": ((), ("out", "stats")),
    "conv3x3_k3_fused": ((), ("out",)),
    "dwconv3x3": ((), ("out",)),
    "stem_im2col": ((), ("out",)),
    "stem_pack": ((1,), ("packed",)),
    "stem_conv": ((), ("out",)),
    "flip_average": ((), ("out",)),
    "decode_final_preds_into": ((3,), ("out",)),
    "normalize_u8": ((), ("out", "packed")),
    "maxpool2x2": ((1,), ("out",)),
    "colstats": ((1, 2), ("sum_out", "sumsq_out", "scratch")),
    "bn_train_fwd": ((4, 5, 6, 7, 8), ()),
    "bn_bwd_reduce": ((3,), ("scratch",)),
    "bn_bwd_apply": ((4,), ("dgamma", "dbeta")),
    "wgrad": ((2,), ()),
    "dwconv3x3_wgrad": ((2,), ()),
    "maxpool2x2_bwd": ((2,), ()),
    "sumpool2x2": ((1,), ()),
    "add_inplace": ((0,), ()),
    "nchw_to_nhwc_bf16_pad": ((1,), ()),
    "small_gemm": ((0,), ()),
    "pack_weights": ((), ("writes",)),          # reads / writes are passed for exactly this bookkeeping
    "jmse_loss_into": ((1, 4), ()),
    "zero_": ((0,), ()),
    "rmsprop_step": ((0, 2), ()),
}
BARRIER = "barrier"      # a call whose footprint is not described: ordered against everything


def _tensors(v):
    if torch.is_tensor(v):
        yield v
    elif isinstance(v, (list, tuple)):
        for x in v:
            if torch.is_tensor(x):
                yield x


def _region(t: torch.Tensor):
    """(storage id, first byte, one past the last byte) touched by a (possibly strided) view."""
    if t.numel() == 0:
        return None
    es = t.element_size()
    lo = t.storage_offset() * es
    span = 1 + sum((s - 1) * abs(st) for s, st in zip(t.shape, t.stride()))
    return (t.untyped_storage().data_ptr(), lo, lo + span * es)


def accesses(name: str, args: Sequence, kwargs: dict):
    """-> (reads, writes) as lists of regions, or BARRIER for a call without a table entry."""
    spec = _WRITES.get(name)
    if spec is None:
        return BARRIER
    wpos, wkw = spec
    reads, writes = [], []
    for i, a in enumerate(args):
        for t in _tensors(a):
            r = _region(t)
            if r is not None:
                (writes if i in wpos else reads).append(r)
    for k, a in kwargs.items():
        for t in _tensors(a):
            r = _region(t)
            if r is not None:
                (writes if k in wkw else reads).append(r)
    return reads, writes


class _RecordingGuard:
    """The recording pass of a plan rebinds the module-global `ops` of hgb200.engine / hgb200.train to a recorder for its
    duration.  That window is process-wide state: one recording at a time (a lock for other threads), and a nested plan
    build from inside the window -- which would record, or launch, through the wrong object -- is refused."""

    def __init__(self):
        import threading
        self._lock = threading.Lock()
        self._owner = None

    def __enter__(self):
        import threading
        me = threading.get_ident()
        if self._owner == me:
            raise RuntimeError("a launch plan is being recorded on this thread: plans cannot be built from inside that pass")
        self._lock.acquire()
        self._owner = me
        return self

    def __exit__(self, *exc):
        self._owner = None
        self._lock.release()
        return False


RECORDING = _RecordingGuard()


class Recorder:
    """Wraps a module of launch wrappers (hgb200.ops) for one eager pass: records, per call, what it reads and
    writes.  `calls` is reset by the caller before each closure."""

    def __init__(self, real):
        self.real = real
        self.calls: List = []

    def __getattr__(self, name):
        fn = getattr(self.real, name)
        if not callable(fn):
            return fn

        def wrapped(*a, **k):
            self.calls.append(accesses(name, a, k))
            return fn(*a, **k)

        return wrapped


class _Storage:
    """Resources (distinct byte ranges) of one storage with their last writer / readers since."""

    def __init__(self):
        self.keys: Dict[Tuple[int, int], int] = {}
        self.lo = np.zeros(0, dtype=np.int64)
        self.hi = np.zeros(0, dtype=np.int64)
        self.over: List[np.ndarray] = []         # per resource: indices of overlapping resources (incl. itself)
        self.writer: List[int] = []
        self.readers: List[List[int]] = []

    def resource(self, lo: int, hi: int) -> int:
        r = self.keys.get((lo, hi))
        if r is not None:
            return r
        r = len(self.writer)
        self.keys[(lo, hi)] = r
        ov = np.nonzero((self.lo < hi) & (self.hi > lo))[0]
        for o in ov:
            self.over[o] = np.append(self.over[o], r)
        self.over.append(np.append(ov, r))
        self.lo = np.append(self.lo, lo)
        self.hi = np.append(self.hi, hi)
        self.writer.append(-1)
        self.readers.append([])
        return r


class LaunchDag:
    def __init__(self, n: int):
        self.n = n
        self.preds: List[List[int]] = [[] for _ in range(n)]
        self.succs: List[List[int]] = [[] for _ in range(n)]
        self.anc: List[int] = [0] * n            # bitset of ancestors

    def check_order(self, order: Sequence[int]) -> bool:
        pos = {op: i for i, op in enumerate(order)}
        return len(pos) == self.n and all(pos[p] < pos[i] for i in range(self.n) for p in self.preds[i])


def build(records: Sequence) -> LaunchDag:
    """records[i]: the `accesses(...)` results of the library calls closure i makes (in order).  An empty list or
    a BARRIER entry makes the closure a barrier."""
    n = len(records)
    dag = LaunchDag(n)
    stores: Dict[int, _Storage] = {}
    last_barrier = -1
    since_barrier: List[int] = []
    for i, calls in enumerate(records):
        reads, writes, barrier = [], [], not calls
        for acc in calls:
            if acc == BARRIER:
                barrier = True
                break
            reads += acc[0]
            writes += acc[1]
        deps = set()
        if barrier:
            deps.update(since_barrier)
            if last_barrier >= 0:
                deps.add(last_barrier)
            # a barrier supersedes every recorded access
            stores.clear()
            last_barrier, since_barrier = i, []
        else:
            if last_barrier >= 0:
                deps.add(last_barrier)
            since_barrier.append(i)
            touched = []
            for (sid, lo, hi) in reads:
                st = stores.setdefault(sid, _Storage())
                r = st.resource(lo, hi)
                for o in st.over[r]:
                    if st.writer[o] >= 0:
                        deps.add(st.writer[o])
                touched.append((st, r, False))
            for (sid, lo, hi) in writes:
                st = stores.setdefault(sid, _Storage())
                r = st.resource(lo, hi)
                for o in st.over[r]:
                    if st.writer[o] >= 0:
                        deps.add(st.writer[o])
                    deps.update(st.readers[o])
                touched.append((st, r, True))
            for st, r, is_write in touched:        # update state after all of this closure's lookups
                if is_write:
                    st.writer[r] = i
                    st.readers[r] = []
                    lo, hi = st.lo[r], st.hi[r]
                    for o in st.over[r]:
                        if o != r and st.lo[o] >= lo and st.hi[o] <= hi:     # fully covered: superseded
                            st.writer[o] = i
                            st.readers[o] = []
                elif not st.readers[r] or st.readers[r][-1] != i:
                    st.readers[r].append(i)
        deps.discard(i)
        # transitive reduction against the ancestors of the other (later) dependencies
        keep, covered = [], 0
        for d in sorted(deps, reverse=True):
            if (covered >> d) & 1:
                continue
            keep.append(d)
            covered |= dag.anc[d] | (1 << d)
        dag.preds[i] = keep
        dag.anc[i] = covered
        for d in keep:
            dag.succs[d].append(i)
    return dag


def assign_streams(dag: LaunchDag, cost: Sequence[float], k: int, leaf_streams: int = 0, leaf_height: float = 0.02):
    """-> (stream of each launch, cross-stream waits of each launch).  cost[i]: estimated duration (any unit).

    leaf_streams > 0 reserves the last `leaf_streams` streams for launches from which every path to a sink is short
    (height <= leaf_height x the critical path: weight-gradient GEMMs, bias sums, the parameter-space chain rule);
    everything else shares the first k - leaf_streams streams, which capture() can give a higher priority so that a
    pending launch of the critical chain is scheduled before a pending leaf."""
    n = dag.n
    height = [0.0] * n                       # longest path to a sink, own cost included
    for i in range(n - 1, -1, -1):
        height[i] = cost[i] + max((height[s] for s in dag.succs[i]), default=0.0)
    if leaf_streams <= 0 or k - leaf_streams < 1:
        leaf_streams = 0
    cut = leaf_height * max(height, default=0.0)
    chain_set, leaf_set = list(range(k - leaf_streams)), list(range(k - leaf_streams, k))
    stream = [0] * n
    waits: List[List[int]] = [[] for _ in range(n)]
    tail = [-1] * k                          # last launch placed on each stream
    avail = [0.0] * k                        # estimated time the stream frees up
    finish = [0.0] * n
    for i in range(n):
        preds = dag.preds[i]
        allowed = leaf_set if (leaf_streams and height[i] <= cut) else chain_set
        ready = max((finish[p] for p in preds), default=0.0)
        choice = None
        # 1. follow a predecessor that is still the tail of its stream -- unless that predecessor has a more
        #    critical successor still to come (leave the chain to it)
        for p in sorted(preds, reverse=True):
            s = stream[p]
            if tail[s] != p or s not in allowed:
                continue
            rivals = [j for j in dag.succs[p] if j > i and height[j] > height[i]]
            if rivals and len(allowed) > 1:
                continue
            choice = s
            break
        if choice is None:
            # 2. a stream whose tail is already an ancestor (or empty): no false edge
            free = [s for s in allowed if tail[s] < 0 or (dag.anc[i] >> tail[s]) & 1]
            if free:
                choice = min(free, key=lambda s: (avail[s], s))
            else:
                # 3. the stream that frees up first
                choice = min(allowed, key=lambda s: (max(avail[s], ready), s))
        stream[i] = choice
        waits[i] = [p for p in preds if stream[p] != choice]
        start = max(ready, avail[choice])
        finish[i] = start + cost[i]
        avail[choice] = finish[i]
        tail[choice] = i
    return stream, waits


def capture(fns: Sequence[Callable], stream_of: Sequence[int], waits: Sequence[Sequence[int]], k: int, device,
            priorities: Optional[Sequence[int]] = None):
    """Capture `fns` into one CUDA graph across k streams.  Without `priorities` stream 0 is the capturing stream;
    with them every stream s is a fresh stream of priority priorities[s] (lower = more urgent, as CUDA counts) and the
    capturing stream only forks and joins.  A closure may enqueue work through another library that is itself
    capture-aware (the NCCL all-reduce nodes of the data-parallel step, hgb200/train.py): whatever it forks rejoins the
    closure's stream before it returns."""
    g = torch.cuda.CUDAGraph()
    need_event = set()
    for w in waits:
        need_event.update(w)
    events: Dict[int, torch.cuda.Event] = {}
    if priorities is None:
        side = [torch.cuda.Stream(device=device) for _ in range(k - 1)]
    else:
        side = [torch.cuda.Stream(device=device, priority=int(priorities[s])) for s in range(k)]
    with torch.cuda.graph(g):
        origin = torch.cuda.current_stream(device)
        streams = ([origin] + side) if priorities is None else side
        start = torch.cuda.Event()
        start.record(origin)
        joined = [st is origin for st in streams]
        for i, fn in enumerate(fns):
            s = stream_of[i]
            st = streams[s]
            if not joined[s]:
                st.wait_event(start)             # fork: brings the side stream into the capture
                joined[s] = True
            for p in waits[i]:
                st.wait_event(events[p])
            with torch.cuda.stream(st):
                fn()
            if i in need_event:
                ev = torch.cuda.Event()
                ev.record(st)
                events[i] = ev
        for s in range(k):                       # join everything back into the origin stream
            if joined[s] and streams[s] is not origin:
                ev = torch.cuda.Event()
                ev.record(streams[s])
                origin.wait_event(ev)
    return g
