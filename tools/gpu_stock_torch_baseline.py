"""Stock PyTorch on the same B200 ("the GPU baseline to beat"): the oracle's restatement of the
reference network (plain torch.nn.functional calls -> cuDNN / cuBLAS kernels, NCHW fp32 exactly as the reference runs
it, plus its most favourable stock settings: TF32, bf16 + channels_last, cudnn.benchmark) timed on the workloads
bench.py measures.  MEASUREMENT SCRIPT, test infrastructure only: run by tools/gpu_profile_r1.sh, never imported by
the product.  Writes one JSON object (path = argv[1], default gpurun_out/stock_torch_b200.json).

    C2: 8-stack J=16 flip-test inference, batch 128 (two forwards per image + the flip average)
    C3: 8-stack training step, batch 32 (train-mode BN, JointsMSE over all stacks, backward, RMSprop)"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hourglass_oracle as H  # noqa: E402
from oracle import train_oracle as T  # noqa: E402

MPII_PERM = [5, 4, 3, 2, 1, 0, 6, 7, 8, 9, 15, 14, 13, 12, 11, 10]


def _time(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _cast(sd, dtype, channels_last):
    out = {}
    for k, v in sd.items():
        v = v.cuda()
        if v.is_floating_point():
            v = v.to(dtype)
        if channels_last and v.dim() == 4:
            v = v.contiguous(memory_format=torch.channels_last)
        out[k] = v
    return out


def inference(sd_cpu, batch, mode):
    dtype = torch.bfloat16 if mode == "bf16_channels_last" else torch.float32
    cl = mode == "bf16_channels_last"
    torch.backends.cudnn.allow_tf32 = mode != "fp32_strict"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32_strict"
    sd = _cast(sd_cpu, dtype, cl)
    x = torch.randn(batch, 3, 256, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(2)).to(dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    perm = torch.tensor(MPII_PERM, device="cuda")

    @torch.no_grad()
    def step():
        hm = H.hg_forward(sd, x)[-1]
        hf = H.hg_forward(sd, x.flip(-1))[-1].flip(-1)[:, perm]
        return 0.5 * (hm + hf)

    ms = _time(step, 3, 5)
    return {"ms_per_step": ms, "images_per_s": batch / ms * 1e3}


def training(sd_cpu, batch, mode):
    torch.backends.cudnn.allow_tf32 = mode != "fp32_strict"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32_strict"
    cl = mode == "autocast_bf16_channels_last"
    sd = _cast(sd_cpu, torch.float32, cl)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if T.is_param(k)}
    work = dict(sd)
    work.update(leaves)
    opt = torch.optim.RMSprop(list(leaves.values()), lr=2.5e-4, momentum=0, weight_decay=0)
    g = torch.Generator("cuda").manual_seed(100)
    x = torch.randn(batch, 3, 256, 256, device="cuda", generator=g)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    target = torch.rand(batch, 16, 64, 64, device="cuda", generator=g)
    tw = (torch.rand(batch, 16, 1, device="cuda", generator=g) < 0.8).float()

    def step():
        H._TRAINING[0] = True
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=cl):
                outs = H.hg_forward(work, x)
            loss = T.joints_mse_torch([o.float() for o in outs], target, tw, True)
        finally:
            H._TRAINING[0] = False
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    ms = _time(step, 3, 5)
    return {"ms_per_step": ms, "images_per_s": batch / ms * 1e3}


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/stock_torch_b200.json"
    torch.backends.cudnn.benchmark = True
    sd = H.make_state_dict(num_stacks=8, num_blocks=1, num_classes=16, seed=0)
    res = {"what": "oracle restatement of the reference network as plain torch ops on cuda:0 (cuDNN/cuBLAS), "
                   "cudnn.benchmark=True, CUDA-event timing, 3 warm-up + 5 timed steps",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "gpu": torch.cuda.get_device_name(0),
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    for mode in ("fp32_strict", "fp32_tf32", "bf16_channels_last"):
        res[f"C2_infer_flip_b128_{mode}"] = inference(sd, 128, mode)
        torch.cuda.empty_cache()
    for mode in ("fp32_strict", "fp32_tf32", "autocast_bf16_channels_last"):
        res[f"C3_train_b32_{mode}"] = training(sd, 32, mode)
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
