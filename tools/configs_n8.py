"""BASELINE.json configs 4 and 5 on N GPUs of one box (run under torchrun; N = WORLD_SIZE, also works at N = 1):

  C4: 8-stack J=17 hourglass, 256x192 inputs, training with target_weight JointsMSE, batch 64 per GPU, data parallel with
      the NCCL all-reduce of the flat gradient buffer (src/runner/trainer.py:82-99 is the step this replaces);
  C5: 4-stack J=21 (hands) and J=14 (CrowdPose) flip-test inference, per-GPU batch 1..128 (global 8..1024 on 8 GPUs),
      batch-sharded, no collective.

Device-timed (CUDA events, barrier + synchronize on both sides, max over ranks); rank 0 prints one JSON line per case.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/configs_n8.py [c4] [c5]"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "hourglass-pose-estimation_b200"))
from bench import Ctx  # noqa: E402


def c4_train(ctx, steps=10, warmup=4):
    import torch.distributed as dist
    from hgb200 import ops
    from hgb200.train import train_engine
    from src.models import hg
    B, J, H, W, lr = 64, 17, 256, 192, 2.5e-4
    torch.manual_seed(0)
    model = hg(num_stacks=8, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum").to(ctx.device).train()
    eng = train_engine(model)
    rng = np.random.RandomState(100 + ctx.rank)
    x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(100 + ctx.rank)).to(ctx.device)
    joints = np.zeros((B, J, 3))
    joints[..., 0], joints[..., 1] = rng.uniform(0, W, (B, J)), rng.uniform(0, H, (B, J))
    vis = (rng.rand(B, J, 1) < 0.8).astype(np.float64).repeat(3, 2)
    jt, vs = torch.from_numpy(joints).to(ctx.device), torch.from_numpy(vis).to(ctx.device)
    def reduce_fn(flat):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)        # issued per bucket inside the step's graph (hgb200/train.py)

    def step():
        mu, wt = ops.joint_centers(jt, vs, (W // 4, H // 4), (W, H), 1)
        tgt = ops.gaussian_target(mu, wt, (W // 4, H // 4), 1)
        return eng.train_step(x, tgt, wt, lr, world_size=ctx.world, all_reduce=reduce_fn if ctx.world > 1 else None)

    for _ in range(warmup):
        loss = step()
    loss0 = float(loss)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
    ops.check_err_word(ctx.device)
    eng.release_graphs()
    return {"config": "C4: COCO 17-joint 8-stack hourglass, 256x192, target_weight loss, training, batch 64/GPU",
            "n_gpus": ctx.world, "images_per_s": ctx.world * B / (ms * 1e-3), "ms_per_step": ms,
            "loss_first_last": [loss0, float(loss)]}


def c5_sweep(ctx, J, batches=(1, 2, 4, 8, 16, 32, 64, 128), steps=10, warmup=3):
    from hgb200 import ops
    from hgb200.infer import FlipTestPipeline
    from src.models import hg
    H = W = 256
    torch.manual_seed(0)
    model = hg(num_stacks=4, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum").to(ctx.device).eval()
    eng = model.engine()
    # a symmetric pairing of the first 2*(J//2) joints: the flip-average kernel only needs SOME permutation table
    pairs = [(2 * i, 2 * i + 1) for i in range(J // 2)]
    out = []
    for B in batches:
        pipe = FlipTestPipeline(eng, B, H, W, flip_pairs=pairs)
        pipe.set_affine(np.tile([[128.0, 128.0]], (B, 1)), np.tile([[1.28, 1.28]], (B, 1)))
        x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(ctx.rank)).to(ctx.device)
        for _ in range(warmup):
            pipe.infer_device(x)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            pipe.infer_device(x)
        e1.record()
        ctx.barrier()
        ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
        ops.check_err_word(ctx.device)
        out.append({"batch_per_gpu": B, "global_batch": B * ctx.world, "ms_per_step": ms,
                    "images_per_s": ctx.world * B / (ms * 1e-3)})
        del pipe
    return {"config": f"C5: 4-stack J={J} flip-test inference sweep, 256x256, batch-sharded, no collective",
            "n_gpus": ctx.world, "sweep": out}


def main():
    what = set(sys.argv[1:]) or {"c4", "c5"}
    ctx = Ctx()
    res = []
    if "c4" in what:
        res.append(c4_train(ctx))
    if "c5" in what:
        b = tuple(int(v) for v in os.environ.get("C5_BATCHES", "1,2,4,8,16,32,64,128").split(","))
        res.append(c5_sweep(ctx, 21, batches=b))
        res.append(c5_sweep(ctx, 14, batches=b))
    if ctx.rank == 0:
        for r in res:
            print(json.dumps(r), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
