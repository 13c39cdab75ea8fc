"""Data-parallel training step on >= 2 GPUs over NCCL (skipped on a one-GPU box; run with `gpurun --gpus 2`).

The step this replaces is the reference's DataParallel loop (src/runner/trainer.py:37,82-99): every replica sees its
shard with its own BatchNorm statistics, the gradients are summed.  Here the sum is an NCCL all-reduce issued per bucket
INSIDE the step's CUDA graph (hgb200/train.py, OVERLAP_ALLREDUCE).  Checked, with the fixed-order reductions on
(HG_DETERMINISTIC: local gradients are bit-reproducible run to run):
  * the exchanged gradient equals the sum over ranks of the gradients each rank computes alone (1-vs-N equality);
  * graph with the exchange inside == eager launch list followed by the same bucket all-reduces, bit for bit;
  * after RMSprop the parameters are bit-identical on every rank."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, os.path.join(REPO, "hourglass-pose-estimation_b200"), HERE):      # the spawned workers import this module too
    if p not in sys.path:
        sys.path.insert(0, p)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import hgb200.train as tr
        from hgb200 import ops
        from hgb200.shard import batch_shard
        from oracle.hourglass_oracle import make_state_dict
        from oracle.make_golden_inputs import train_inputs
        from src.models import hg
        tr.DETERMINISTIC = True
        tr.OVERLAP_ALLREDUCE = True
        dev = torch.device("cuda", rank)
        S, J, B, H, W = 2, 16, 8, 128, 128
        sd = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0)
        model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum")
        model.load_state_dict(sd)
        model = model.to(dev).train()
        eng = tr.train_engine(model)
        x, tg, tw = train_inputs(1, B, J, H, W, 1)[0]
        a, b = batch_shard(B, world, rank)
        xs, ts, ws = x[a:b].to(dev), tg[a:b].to(dev), tw[a:b].to(dev)
        calls = []

        def reduce_fn(flat):
            calls.append(flat.numel())
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)

        G, n = eng.store.G, eng.store.count
        # (1) alone: lr = 0 leaves the parameters where they are
        eng.train_step(xs, ts, ws, 0.0, world_size=world)
        local = G[:n].clone()
        # (2) eager launch list + bucket all-reduces
        eng.train_step(xs, ts, ws, 0.0, use_graph=False, world_size=world, all_reduce=reduce_fn)
        eager = G[:n].clone()
        n_eager_calls = len(calls)
        # (3) the graph with the exchange inside, twice (capture + replay)
        for _ in range(2):
            eng.train_step(xs, ts, ws, 0.0, world_size=world, all_reduce=reduce_fn)
        graph = G[:n].clone()
        torch.cuda.synchronize(dev)
        ops.check_err_word(dev)
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        want = torch.stack(gathered).double().sum(0)
        scale = float(want.abs().max())
        err = float((graph.double() - want).abs().max()) / scale
        # (4) one real step: parameters identical everywhere afterwards (the second moments restart from zero: the
        #     local step of (1) fed them with per-rank gradients)
        eng.store.V.zero_()
        eng.train_step(xs, ts, ws, 2.5e-4, world_size=world, all_reduce=reduce_fn)
        P = eng.store.P[:n].clone()
        allP = [torch.empty_like(P) for _ in range(world)]
        dist.all_gather(allP, P)
        if rank == 0:
            ret["err"] = err
            ret["graph_equals_eager"] = bool(torch.equal(graph, eager))
            ret["buckets"] = len(eng.grad_buckets())
            ret["eager_calls"] = n_eager_calls - 1          # minus the communicator warm-up
            ret["params_identical"] = all(torch.equal(allP[0], t) for t in allP)
            ret["moved"] = float((P - eng.store.P[:n]).abs().max()) == 0.0 and float(graph.abs().max()) > 0
        eng.release_graphs()                                # graphs with NCCL nodes go before the communicator
    finally:
        dist.destroy_process_group()


def test_in_graph_bucketed_allreduce_matches_the_sum_of_local_gradients():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    print("\nNCCL world 2:", dict(ret))
    assert ret["buckets"] == 2 + 2 and ret["eager_calls"] == ret["buckets"]
    assert ret["err"] <= 1e-6, ret["err"]                   # fp32 sum of two addends: exact up to the reduction's own rounding
    assert ret["graph_equals_eager"]
    assert ret["params_identical"] and ret["moved"]
