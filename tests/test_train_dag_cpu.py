"""Launch DAG of the training step (hgb200/dag.py) on CPU: dependency analysis from byte ranges, stream
assignment, and -- the real check -- the whole step executed in RANDOM topological orders of the DAG through
the CPU emulation must give bit-identical gradients, heat maps, loss and BN statistics to the sequential order
(any missing read-after-write / write-after-read edge, e.g. through a recycled gradient buffer, shows up)."""
import os
import random
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fake_ops  # noqa: E402
from hgb200 import dag  # noqa: E402
from oracle.hourglass_oracle import make_state_dict  # noqa: E402
from oracle.make_golden_inputs import train_inputs  # noqa: E402


def _acc(reads=(), writes=()):
    return [([("s", lo, hi) for lo, hi in reads], [("s", lo, hi) for lo, hi in writes])]


def test_edges_from_byte_ranges():
    recs = [
        _acc(writes=[(0, 100)]),                 # 0 writes A
        _acc(reads=[(0, 50)]),                   # 1 reads half of A            -> RAW on 0
        _acc(reads=[(50, 100)]),                 # 2 reads the other half       -> RAW on 0
        _acc(writes=[(0, 50)]),                  # 3 overwrites the first half  -> WAR on 1 (WAW on 0 is implied)
        _acc(reads=[(0, 100)]),                  # 4 reads all                  -> RAW on 3 (0 implied)
        _acc(writes=[(200, 300)]),               # 5 unrelated
        [],                                      # 6 barrier
        _acc(reads=[(200, 300)]),                # 7 after the barrier
    ]
    d = dag.build(recs)
    assert d.preds[0] == [] and d.preds[1] == [0] and d.preds[2] == [0]
    assert d.preds[3] == [1]
    assert sorted(d.preds[4]) == [3]
    assert d.preds[5] == []
    assert sorted(d.preds[6]) == [2, 4, 5]
    assert d.preds[7] == [6]
    assert d.check_order(range(8)) and not d.check_order([1, 0, 2, 3, 4, 5, 6, 7])


def test_every_launch_wrapper_the_plans_use_has_a_footprint_entry():
    """A wrapper without a _WRITES entry is a BARRIER in the launch DAG (correct but serialising); one whose entry forgets
    an output is a missed dependency.  The fused bottleneck tail: reads the halo buffer, weights, residual and the
    low-resolution operand, writes `out` only."""
    t = lambda *shape: torch.zeros(*shape)
    halo, w2, b2, w3, b3 = t(100), t(128, 1152), t(128), t(256, 128), t(256)
    res, up, out = t(2, 4, 4, 256), t(2, 2, 2, 256), t(2, 4, 4, 256)
    reads, writes = dag.accesses("conv3x3_k3_fused", (halo, w2, b2, w3, b3), dict(n=2, h=4, w=4, residual=res, up_low=up, out=out))
    region = lambda x: (x.untyped_storage().data_ptr(), 0, x.numel() * 4)
    assert writes == [region(out)]
    assert {region(x) for x in (halo, w2, b2, w3, b3, res, up)} == set(reads)
    for name in ("conv_nhwc", "conv3x3_halo", "conv3x3_k3_fused", "maxpool2x2", "dwconv3x3", "stem_pack", "stem_conv",
                 "flip_average", "decode_final_preds_into"):
        assert name in dag._WRITES, name


def test_stream_assignment_covers_every_edge():
    rnd = random.Random(0)
    recs = []
    for i in range(300):
        reads = [(64 * rnd.randrange(40), 64 * rnd.randrange(40) + 64) for _ in range(2)]
        reads = [(lo, max(hi, lo + 64)) for lo, hi in reads]
        lo = 64 * rnd.randrange(40)
        recs.append(_acc(reads=reads, writes=[(lo, lo + 64)]))
    d = dag.build(recs)
    for k in (1, 2, 5):
        stream_of, waits = dag.assign_streams(d, [1.0] * d.n, k)
        last = {}
        for i in range(d.n):
            for p in d.preds[i]:
                assert stream_of[p] == stream_of[i] or p in waits[i]
            assert 0 <= stream_of[i] < k
            last[stream_of[i]] = i
        if k == 1:
            assert not any(waits)


@pytest.fixture()
def cpu_train(monkeypatch):
    import hgb200.train as tr
    monkeypatch.setattr(tr, "ops", fake_ops)
    monkeypatch.setattr(tr, "STREAMS", 6)
    return tr


def _random_topological_order(d, rnd):
    indeg = [len(p) for p in d.preds]
    ready = [i for i in range(d.n) if indeg[i] == 0]
    order = []
    while ready:
        i = ready.pop(rnd.randrange(len(ready)))
        order.append(i)
        for s in d.succs[i]:
            indeg[s] -= 1
            if indeg[s] == 0:
                ready.append(s)
    assert len(order) == d.n
    return order


@pytest.mark.parametrize("S,J,B,H,W", [(2, 16, 2, 64, 64), (1, 17, 1, 64, 128)])
def test_random_topological_orders_give_identical_steps(cpu_train, S, J, B, H, W):
    from src.models import hg
    sd = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0)
    model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum")
    model.load_state_dict(sd)
    model.train()
    eng = cpu_train.TrainEngine(model, "cpu")
    plan = eng.plan_for(B, H, W)
    x, tg, tw = train_inputs(1, B, J, H, W, 1)[0]
    plan.input.copy_(x)
    plan.target.copy_(tg)
    plan.target_weight.copy_(tw.reshape(B, J))
    bns = [(b.rm, b.rv, b.nbt) for b in eng._bns]
    keep = [tuple(t.clone() for t in b) for b in bns]

    def snapshot():
        return [eng.store.G.clone(), plan.loss.clone()] + [o.clone() for o in plan.outputs] + \
               [t.clone() for b in bns for t in b]

    def restore():
        for b, k in zip(bns, keep):
            for t, kt in zip(b, k):
                t.copy_(kt)

    fns = plan.launches("step")
    for fn in fns:
        fn()
    want = snapshot()
    d, stream_of, waits = plan.schedule("step")
    assert d.n == len(fns) == len(plan.meta)
    assert len(set(stream_of)) > 1                      # the step really has parallel branches
    # weight gradients are leaves: most launches of the backward pass are off the critical chain
    depth = [0] * d.n
    for i in range(d.n):
        depth[i] = 1 + max((depth[p] for p in d.preds[i]), default=0)
    assert max(depth) < 0.8 * d.n
    for seed in range(3):
        restore()
        order = _random_topological_order(d, random.Random(seed))
        assert d.check_order(order)
        for i in order:
            fns[i]()
        got = snapshot()
        for a, b in zip(got, want):
            assert torch.equal(a, b)
    # the separately captured halves (autograd drop-in) carry their own DAGs
    for which in ("fwd", "bwd"):
        dd, so, _ = plan.schedule(which)
        assert dd.n == len(plan.launches(which))


def test_gradient_buckets_and_their_place_in_the_launch_dag(monkeypatch):
    """Data-parallel step (hgb200/train.py, OVERLAP_ALLREDUCE): the flat gradient buffer is cut into buckets that the
    backward pass completes one after the other (hourglass of the last stack first), and every all-reduce node of the
    step's graph waits for exactly the launches that last wrote its slice: each launch that writes into the bucket must be
    an ancestor (or one) of the node's waits, and no wait may be a launch that does not touch the bucket's history."""
    import hgb200.train as tr
    from src.models import hg
    monkeypatch.setattr(tr, "ops", fake_ops)
    S, J, B, H, W = 3, 16, 2, 64, 64
    model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum")
    model.load_state_dict(make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0))
    model.train()
    eng = tr.TrainEngine(model, "cpu")
    buckets = eng.grad_buckets()
    st = eng.store
    # a partition of [0, count): disjoint, complete; the first S buckets are hg.<S-1> ... hg.<0>
    cover = sorted(buckets)
    assert cover[0][0] == 0 and cover[-1][1] == st.count
    assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    # order: hg.<S-1> ... hg.<1>, the block behind the hourglasses, hg.<0>, the block in front of them (stem, layer1-3)
    where = {i: (j if i > 0 else S) for j, i in enumerate(range(S - 1, -1, -1))}
    for i, j in where.items():
        names = [n for n in st.slots if n.startswith(f"hg.{i}.")]
        lo = min(st.slots[n][0] for n in names)
        hi = max(st.slots[n][0] + st.slots[n][1] for n in names)
        assert buckets[j][0] == lo and hi <= buckets[j][1] <= hi + 3, (j, buckets[j], lo, hi)
    assert len(buckets) == S + 2 and buckets[-1][0] == 0 and buckets[S - 1][1] == st.count
    assert buckets[-1][1] - buckets[-1][0] < 0.06 * st.count                # what no launch can hide: stem + layer1-3
    plan = eng.plan_for(B, H, W)
    stream, waits = plan.comm_schedule(buckets)
    assert stream == tr.STREAMS and len(waits) == len(buckets)
    d, _, _ = plan.schedule("step")
    gptr = st.G.untyped_storage().data_ptr()
    anc_cache = {}

    def ancestors(i):
        if i not in anc_cache:
            s = {i}
            for p in d.preds[i]:
                s |= ancestors(p)
            anc_cache[i] = s
        return anc_cache[i]

    first_wait = []
    for (lo, hi), w in zip(buckets, waits):
        assert w, "every bucket has a writer"
        closure = set()
        for p in w:
            closure |= ancestors(p)
        writers = [i for i, rec in enumerate(plan.records)
                   if any(acc != dag.BARRIER and any(r[0] == gptr and r[1] < hi * 4 and r[2] > lo * 4 for r in acc[1]) for acc in rec)]
        assert writers and set(writers) <= closure
        assert set(w) <= set(writers)                  # a node waits only for launches that write its own slice
        first_wait.append(max(w))
    # the buckets complete in the order they are issued on the exchange stream
    assert first_wait == sorted(first_wait), first_wait


def test_recording_window_is_exclusive():
    """The recording pass rebinds hgb200.train.ops / hgb200.engine.ops process-wide (dag.RECORDING): a nested plan build on
    the same thread is refused instead of recording through the wrong object, and the guard is free again afterwards."""
    with dag.RECORDING:
        with pytest.raises(RuntimeError):
            with dag.RECORDING:
                pass
    with dag.RECORDING:
        pass
