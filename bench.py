#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 stacked-hourglass hot path.

BASELINE.json's metric has two halves ("images/sec 8-stack HG 256x256 fwd/train at 1/2/4/8 B200"); ONE run measures both
and prints ONE JSON line:

  headline (value / e2e / roofline / clocks)   C2 = BASELINE.json configs[1]: 8-stack hourglass, MPII 16 joints, 256x256,
                                               batch 128 per GPU, flip-test inference (two forwards per image),
                                               flip-average, arg-max + quarter-pixel + affine decode.  Batch-sharded over
                                               the GPUs, no collective.
  "train": {...} sub-record                    C3 = configs[2]: 8-stack training step, batch 32 per GPU, on-device Gaussian
                                               targets, JointsMSE over all stacks, backward, NCCL all-reduce of the flat
                                               gradient buffer (N>1), fused RMSprop.  Carries its own value, ms_per_step,
                                               e2e, allreduce_ms, roofline, tensor_frac, loss_first_last and clocks.

  python bench.py [--gpus N] [--steps K] [--warmup W]            both workloads (default)
  python bench.py --workload infer|train ...                     one of them (the train record is then the line itself)
  python bench.py --impl reference ...                           the reference's own code on the host CPUs: the unmodified
                                                                 reference modules vendored by oracle/vendor_reference.py
                                                                 (oracle/_ref, kind "reference"), else the oracle port

N>1 is launched by torchrun (one rank per GPU).  Metric: images/sec (original images, whole job over all GPUs); one "step"
= one batch per GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "hourglass-pose-estimation_b200")
for _p in (REPO, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

FWD_GFLOP_PER_IMAGE = 56.189          # 2*MAC over the reference's Conv2d layers, 8-stack J=16 256x256
TRAIN_GFLOP_PER_IMAGE = 3 * FWD_GFLOP_PER_IMAGE      # fwd + dgrad + wgrad
METRIC = "images/sec 8-stack HG 256x256 flip-test inference"
WORKLOAD = "C2: 8-stack hourglass MPII 16-joint inference with flip test, batch 128, bf16"
TRAIN_METRIC = "images/sec 8-stack HG 256x256 train"
TRAIN_WORKLOAD = "C3: 8-stack hourglass MPII training, intermediate-supervision JointsMSELoss, RMSprop, batch 32/GPU, bf16"
MPII_MEAN, MPII_STD = (0.4327, 0.4440, 0.4404), (0.2468, 0.2410, 0.2458)     # reference: src/runner/estimator.py:43-44
MPII_PERM = [5, 4, 3, 2, 1, 0, 6, 7, 8, 9, 15, 14, 13, 12, 11, 10]


def infer_config(world: int, B: int) -> dict:
    return {"workload": WORKLOAD, "images_per_gpu_per_step": B, "forwards_per_image": 2,
            "parallelism": f"batch-sharded x{world}, no collective",
            "l2": "working set (>3 GB of activations per step) far exceeds the 126 MB L2; no flush needed",
            "e2e_path": "FlipTestPipeline.infer_host_u8: pinned uint8 HWC frames -> H2D (double-buffered) -> "
                        "hg_normalize_u8_nhwc -> graph -> D2H fp64 coordinates",
            "note": "value (device-resident fp32 input) and e2e are timed in separate back-to-back regions; under "
                    "sw_power_cap the SM clock drifts a few percent between them",
            "weights": "random init (torch default), randomised BN statistics"}


def train_config(world: int, B: int) -> dict:
    return {"workload": TRAIN_WORKLOAD, "images_per_gpu_per_step": B, "global_batch": B * world,
            "parallelism": (f"data parallel x{world}, one NCCL all-reduce (sum) of the flat fp32 gradient buffer per step"
                            if world > 1 else "single GPU, no collective"),
            "l2": "working set (>9 GB of saved activations per step) far exceeds the 126 MB L2; no flush needed",
            "weights": "random init (torch default)"}


BOTH = "  [headline value/e2e/roofline]  +  "


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "25"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def started(self) -> bool:
        """True once nvidia-smi has written its first sample (it needs ~0.1-0.3 s to start): the caller keeps the GPU
        under the benchmark's own load with untimed steps until then, so the timed region is never missed."""
        if self.p is None:
            return True
        try:
            return os.path.getsize(self.f.name) > 0
        except OSError:
            return True

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def ncu_traffic(workload: str, kernel_class: str, images: int):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full captures
    (profiles/r2_ncu_traffic.json, else round 1's), or None when no capture matches this kernel class and batch."""
    for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        try:
            with open(os.path.join(REPO, "profiles", name)) as f:
                ent = json.load(f).get(workload, {}).get(f"{kernel_class}@{images}")
            if ent:
                return ent["dram_bytes"]
        except (OSError, ValueError, KeyError):
            continue
    return None


def randomise_bn(model, seed):
    """Randomised BN statistics / affine so that folding is non-trivial."""
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.weight.copy_(0.75 + 0.5 * torch.rand(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))
    return model


def build_model(device, seed=0):
    """Random-init weights of the named architecture (torch default init, as the reference's constructor does)."""
    import torch
    from src.models import hg
    torch.manual_seed(seed)
    model = hg(num_stacks=8, num_blocks=1, num_classes=16, mobile=False, skip_mode="sum", out_res=64)
    return randomise_bn(model, seed + 1).to(device).eval()


# ------------------------------------------------------------------------------------------------ the reference's own code
_REF = {}


def import_reference():
    """The vendored, unmodified reference package (oracle/_ref/src: models, loss, utils) as a dict of modules, or None.
    It is imported under its own name `src` with this repo's `src` package parked aside, then parked itself, so both can
    live in one process."""
    if "mods" in _REF:
        return _REF["mods"]
    _REF["mods"] = None
    ref_root = os.path.join(REPO, "oracle", "_ref")
    try:
        from oracle.vendor_reference import verify
        if not verify():
            return None
    except Exception:
        return None

    def is_src(k):
        return k == "src" or k.startswith("src.")

    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if is_src(k)}
    sys.path.insert(0, ref_root)
    try:
        import importlib
        for name in ("src.models", "src.loss.mse", "src.utils.inference", "src.utils.evaluation", "src.utils.transforms"):
            importlib.import_module(name)
        _REF["mods"] = {k: v for k, v in sys.modules.items() if is_src(k)}
    except Exception as e:                                   # e.g. cv2 missing on this host
        _REF["error"] = f"{type(e).__name__}: {e}"
    finally:
        sys.path.remove(ref_root)
        for k in list(sys.modules):
            if is_src(k):
                del sys.modules[k]
        sys.modules.update(saved)
    return _REF["mods"]


def reference_model(ref, seed=0):
    import torch
    torch.manual_seed(seed)
    m = ref["src.models"].hg(num_stacks=8, num_blocks=1, num_classes=16, mobile=False, skip_mode="sum", out_res=64)
    return randomise_bn(m, seed + 1)


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_reference_rate(n_images: int, steps: int, warmup: int):
    """The reference's flip-test inference on the host cores for `n_images` per step: fp32 NCHW forward of the 8-stack
    network on the batch and its mirror image, the flip average (SURVEY A12's 3-line definition) and get_final_preds_v1
    for every image.  Uses the unmodified reference modules when oracle/_ref is present, else the oracle port."""
    import numpy as np
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = torch.randn(n_images, 3, 256, 256, generator=torch.Generator().manual_seed(2))
    ref = import_reference()
    if ref is not None:
        kind = "reference"
        model = reference_model(ref).eval()
        final_preds = ref["src.utils.inference"].get_final_preds_v1
        perm = torch.tensor(MPII_PERM)
        center, scale = np.array([128.0, 128.0]), np.array([1.28, 1.28])

        def step():
            with torch.no_grad():
                hm = model(x)[-1]
                hf = model(x.flip(-1))[-1].flip(-1)[:, perm]
                avg = 0.5 * (hm + hf)
            # the reference decodes batch element 0 only (inference.py:49,55): one call per image
            return [final_preds(avg[i:i + 1], center, scale, (64, 64)) for i in range(n_images)]
    else:
        kind = "port"
        from oracle.hourglass_oracle import make_state_dict, hg_forward
        from oracle import decode_oracle as D
        sd = make_state_dict(num_stacks=8, num_blocks=1, num_classes=16, seed=0)
        centers = np.tile(np.array([[128.0, 128.0]]), (n_images, 1))
        scales = np.tile(np.array([[1.28, 1.28]]), (n_images, 1))

        def step():
            with torch.no_grad():
                hm = hg_forward(sd, x)[-1].numpy()
                hf = hg_forward(sd, x.flip(-1))[-1].numpy()
            avg = D.flip_average(hm, hf, D.MPII_FLIP_PAIRS)
            return D.get_final_preds_batch(avg, centers, scales, (64, 64))

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_images / dt, dt * 1e3, cores, kind


def cpu_train_rate(n_images: int, steps: int, warmup: int = 1):
    """The reference's training step (trainer.py:82-99) on the host cores on a bounded sample: train-mode forward, MSELoss
    over all stacks, backward, RMSprop.  Unmodified reference modules when oracle/_ref is present, else the oracle port."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle.make_golden_inputs import train_inputs
    warmup = max(1, warmup)
    batches = train_inputs(1, n_images, 16, 256, 256, 2)
    ref = import_reference()
    if ref is not None:
        kind = "reference"
        model = reference_model(ref).train()
        crit = ref["src.loss.mse"].MSELoss(use_target_weight=True)
        opt = torch.optim.RMSprop(model.parameters(), lr=2.5e-4, momentum=0, weight_decay=0)      # trainer.py:39-41

        def step(b):
            x, tgt, tw = b
            loss = crit(model(x), tgt, tw)
            opt.zero_grad()
            loss.backward()
            opt.step()
    else:
        kind = "port"
        from oracle.hourglass_oracle import make_state_dict
        from oracle import train_oracle as T
        sd = make_state_dict(num_stacks=8, num_blocks=1, num_classes=16, seed=0)
        state = {}

        def step(b):
            _, _, grads = T.forward_backward(sd, *b)
            T.rmsprop_update(sd, grads, state, 2.5e-4)

    for _ in range(warmup):
        step(batches[0])
    t0 = time.perf_counter()
    for i in range(steps):
        step(batches[(i + 1) & 1])
    dt = (time.perf_counter() - t0) / steps
    return n_images / dt, dt * 1e3, cores, kind


def run_reference(args):
    """--impl reference: rank 0 alone times the reference's CPU path -- every host thread, this arm's metric / config, K
    timed steps after W warm-up steps, each step a bounded sample of the workload (4 of the 128 images; 2 of the 32)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    line = None
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.workload in ("all", "infer"):
        n_img = 4
        rate, ms, cores, kind = cpu_reference_rate(n_img, steps, warmup)
        sample = (f"{n_img} images/step (of the {args.batch}-image batch), flip-test forward + flip average + "
                  f"get_final_preds_v1, fp32, {cores} threads")
        line = {
            "impl": "reference", "metric": METRIC, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": infer_config(args.gpus, args.batch),
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
    if args.workload in ("all", "train"):
        n_img = 2
        rate, ms, cores, kind = cpu_train_rate(n_img, steps, warmup)
        sample = (f"{n_img} images/step (of the {args.train_batch}-image batch), 8-stack train step fwd+bwd+RMSprop, fp32, "
                  f"{cores} threads")
        tr = {"impl": "reference", "metric": TRAIN_METRIC, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
              "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": train_config(args.gpus, args.train_batch),
              "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
              "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0}
        if line is None:
            line = tr
        else:
            line["train"] = tr
            line["config"]["workload"] = WORKLOAD + BOTH + TRAIN_WORKLOAD + "  ['train' sub-record]"
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ stock torch on the same GPU
def gpu_baseline(device, do_infer: bool, do_train: bool, steps: int = 3):
    """SURVEY 8(d) "the GPU baseline to beat": the reference network executed by STOCK PyTorch (cuDNN / cuBLAS,
    cudnn.benchmark=True as trainer.py:36 sets it) on this B200 in its most favourable stock setting (bf16 + channels_last;
    autocast for training), on the same workloads.  The reference's own nn.Module when oracle/_ref is present, else the
    oracle's functional restatement.  Not the product: nothing here touches libhgb200."""
    import torch
    out = {"what": "stock PyTorch (cuDNN/cuBLAS) on the same GPU, bf16 + channels_last, cudnn.benchmark, CUDA events, "
                   f"2 warm-up + {steps} timed steps", "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    ref = import_reference()
    out["kind"] = "reference" if ref is not None else "port"

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / steps

    try:
        if ref is not None:
            model = reference_model(ref).to(device)
        else:
            from oracle import hourglass_oracle as H
            from oracle import train_oracle as T
            sd32 = {k: v.to(device) for k, v in H.make_state_dict(num_stacks=8, num_blocks=1, num_classes=16, seed=0).items()}
        g = torch.Generator(device=device).manual_seed(2)
        if do_infer:
            B = 128
            x = torch.randn(B, 3, 256, 256, device=device, generator=g).to(torch.bfloat16).contiguous(
                memory_format=torch.channels_last)
            perm = torch.tensor(MPII_PERM, device=device)
            if ref is not None:
                m16 = model.eval().to(torch.bfloat16).to(memory_format=torch.channels_last)
                fwd = lambda t: m16(t)[-1]                                                  # noqa: E731
            else:
                sd16 = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v) for k, v in sd32.items()}
                sd16 = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd16.items()}
                fwd = lambda t: H.hg_forward(sd16, t)[-1]                                   # noqa: E731

            @torch.no_grad()
            def step():
                hm = fwd(x)
                hf = fwd(x.flip(-1)).flip(-1)[:, perm]
                return 0.5 * (hm + hf)

            ms = timed(step)
            out["infer"] = {"value": B / ms * 1e3, "unit": "images/s", "ms_per_step": ms, "batch": B}
            del x
            if ref is not None:
                model = model.float()
            torch.cuda.empty_cache()
        if do_train:
            B = 32
            x = torch.randn(B, 3, 256, 256, device=device, generator=g).contiguous(memory_format=torch.channels_last)
            target = torch.rand(B, 16, 64, 64, device=device, generator=g)
            tw = (torch.rand(B, 16, 1, device=device, generator=g) < 0.8).float()
            if ref is not None:
                m = model.train().to(memory_format=torch.channels_last)
                crit = ref["src.loss.mse"].MSELoss(use_target_weight=True)
                opt = torch.optim.RMSprop(m.parameters(), lr=2.5e-4, momentum=0, weight_decay=0)

                def step():
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        outs = m(x)
                    loss = crit([o.float() for o in outs], target, tw)
                    opt.zero_grad(set_to_none=True)
                    loss.backward()
                    opt.step()
            else:
                leaves = {k: v.clone().requires_grad_(True) for k, v in sd32.items() if T.is_param(k)}
                work = dict(sd32)
                work.update(leaves)
                opt = torch.optim.RMSprop(list(leaves.values()), lr=2.5e-4, momentum=0, weight_decay=0)

                def step():
                    H._TRAINING[0] = True
                    try:
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            outs = H.hg_forward(work, x)
                        loss = T.joints_mse_torch([o.float() for o in outs], target, tw, True)
                    finally:
                        H._TRAINING[0] = False
                    opt.zero_grad(set_to_none=True)
                    loss.backward()
                    opt.step()

            ms = timed(step)
            out["train"] = {"value": B / ms * 1e3, "unit": "images/s", "ms_per_step": ms, "batch": B}
    except Exception as e:                                   # a baseline must never take the benchmark down
        out["error"] = f"{type(e).__name__}: {e}"
    finally:
        torch.backends.cudnn.benchmark = prev
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ helpers
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the sm_100a path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.device)
        from hgb200 import lib
        lib.check(lib.hg_check_device(), "hg_check_device")

    def barrier(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier()
            torch.cuda.synchronize(self.device)

    def max_over_ranks(self, v: float) -> float:
        import torch
        import torch.distributed as dist
        t = torch.tensor([v], dtype=torch.float64, device=self.device)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.destroy_process_group()


def class_table(per_launch, meta):
    classes = {}
    for ms, m in zip(per_launch, meta):
        c = classes.setdefault(m["op"], dict(ms=0.0, n=0, flops=m["flops"], bytes=m["bytes"], kind=m["kind"]))
        c["ms"] += ms
        c["n"] += 1
    return classes


def roofline_of(top, avg_ms, hbm_peak, tf_sustained):
    if top["kind"] == "conv" and top["flops"] / max(top["bytes"], 1) > tf_sustained * 1e12 / (hbm_peak * 1e9):
        r = {"bound": "tensor", "achieved": top["flops"] / (avg_ms * 1e-3) / 1e12, "peak": tf_sustained, "unit": "TFLOP/s"}
    else:
        r = {"bound": "hbm", "achieved": top["bytes"] / (avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"}
    r["frac"] = r["achieved"] / r["peak"]
    return r


def write_breakdown(path, header_lines, classes, total_ms):
    with open(path, "w") as f:
        for h in header_lines:
            f.write(h + "\n")
        f.write("class,launches,total_ms,avg_ms,share,TFLOP/s,GB/s\n")
        for name, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"]):
            a = c["ms"] / c["n"]
            f.write(f"{name},{c['n']},{c['ms']:.4f},{a:.4f},{c['ms']/total_ms:.4f},"
                    f"{c['flops']/(a*1e-3)/1e12:.1f},{c['bytes']/(a*1e-3)/1e9:.0f}\n")


# ------------------------------------------------------------------------------------------------ C2: flip-test inference
def bench_infer(ctx: Ctx, args):
    import numpy as np
    import torch
    from hgb200 import ops
    from hgb200.infer import FlipTestPipeline
    device, world, rank = ctx.device, ctx.world, ctx.rank
    steps, warmup = args.steps, max(3, args.warmup)
    B = args.batch
    model = build_model(device)
    engine = model.engine(device)
    pipe = FlipTestPipeline(engine, B, 256, 256)
    pipe.set_affine(np.tile([[128.0, 128.0]], (B, 1)), np.tile([[1.28, 1.28]], (B, 1)))
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    # the deployment input: uint8 HWC crops (what the dataset's warpAffine / the estimator's frame grab leave on the host);
    # ToTensor + Normalize run on the device (hg_normalize_u8_nhwc), so a quarter of the fp32 bytes cross PCIe
    host_u8 = [torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(2)]
    x_dev = ops.normalize_u8(host_u8[0].to(device), MPII_MEAN, MPII_STD)

    # ---------------- device-resident throughput (inputs already in HBM) ----------------
    for _ in range(warmup):
        pipe.infer_device(x_dev)
    ctx.barrier()
    ops.check_err_word(device)
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    t_wait = time.perf_counter()
    while sampler is not None and not sampler.started() and time.perf_counter() - t_wait < 3.0:
        pipe.infer_device(x_dev)                       # untimed: same load while nvidia-smi starts up
        torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(steps):
        pipe.infer_device(x_dev)
    e1.record()
    ctx.barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    ms_step = ctx.max_over_ranks(ms_total) / steps
    value = world * B / (ms_step * 1e-3)
    ops.check_err_word(device)

    # ---------------- end to end: pinned uint8 host frames -> H2D -> normalise -> graph -> D2H coordinates ----------------
    def host_batches(k):
        for i in range(k):
            yield host_u8[i & 1]

    for _ in pipe.infer_host_u8(host_batches(2), MPII_MEAN, MPII_STD):
        pass
    ctx.barrier()
    t0 = time.perf_counter()
    n_out = 0
    for coords in pipe.infer_host_u8(host_batches(steps), MPII_MEAN, MPII_STD):
        n_out += coords.shape[0]
    torch.cuda.synchronize(device)
    t_e2e = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * steps / t_e2e
    assert n_out == B * steps
    ops.check_err_word(device)

    # ---------------- roofline of the dominant kernel (CUDA events around every launch, eager replay) ----------
    hbm_peak, tf_burst, tf_sustained, peak_src = load_peaks()
    per_launch = pipe.plan.profile(iters=2)
    classes = class_table(per_launch, pipe.plan.meta)
    total_ms = sum(per_launch)
    top_name, top = max(classes.items(), key=lambda kv: kv[1]["ms"])
    avg_ms = top["ms"] / top["n"]
    roofline = roofline_of(top, avg_ms, hbm_peak, tf_sustained)
    roofline.update(traffic=ncu_traffic("infer", top_name, B), kernel=top_name, launches_per_step=top["n"],
                    share_of_step=top["ms"] / total_ms,
                    peak_source=f"{peak_src} (sustained bf16 / copy bandwidth, MEASURED_PEAKS.json)")
    if args.breakdown and rank == 0:
        write_breakdown(args.breakdown, [f"# per-kernel-class device time, eager replay with CUDA events, batch {B} (x2 flip) ; "
                                         f"total {total_ms:.3f} ms ; graph step {ms_step:.3f} ms"], classes, total_ms)
    launches = pipe.launches_per_batch
    del pipe, engine, model
    torch.cuda.empty_cache()
    flops_per_step = 2 * FWD_GFLOP_PER_IMAGE * 1e9 * B     # two forwards per image
    tf = flops_per_step / (ms_step * 1e-3) / 1e12
    return {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": infer_config(world, B),
        "tensor_tflops": tf, "tensor_frac_of_measured_peak": tf / tf_sustained,
        "tensor_frac": {"of_sustained": tf / tf_sustained, "of_burst": tf / tf_burst, "of_nominal_2250": tf / 2250.0},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 256 * 256 * 3,
                "d2h_bytes_per_step": B * 16 * 2 * 8},
        "gpu_launches": (launches + 1) * steps,
        "clocks": clocks,
    }


# ------------------------------------------------------------------------------------------------ C3: training step
def bench_train(ctx: Ctx, args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from hgb200 import ops
    from hgb200.train import train_engine
    from hgb200.prefetch import DevicePrefetcher, LaggedScalar
    import hgb200.train as _tr
    from src.models import hg
    device, world, rank = ctx.device, ctx.world, ctx.rank
    steps, warmup = args.steps, max(3, args.warmup)
    B = args.train_batch
    J, H, W, lr = 16, 256, 256, 2.5e-4
    torch.manual_seed(0)                                     # identical initial weights on every rank
    model = hg(num_stacks=8, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum", out_res=64).to(device).train()
    eng = train_engine(model)
    rng = np.random.RandomState(100 + rank)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    host_x = [torch.randn(B, 3, H, W, generator=g).pin_memory() for _ in range(2)]
    joints = np.zeros((B, J, 3))
    joints[..., 0], joints[..., 1] = rng.uniform(0, W, (B, J)), rng.uniform(0, H, (B, J))
    vis = (rng.rand(B, J, 1) < 0.8).astype(np.float64).repeat(3, 2)
    host_j, host_v = torch.from_numpy(joints).pin_memory(), torch.from_numpy(vis).pin_memory()
    x_dev, j_dev, v_dev = host_x[0].to(device), host_j.to(device), host_v.to(device)
    def reduce_fn(flat):
        """The step's exchange: NCCL all-reduce (sum) of a slice of the flat fp32 gradient buffer.  hgb200.train issues it
        per bucket INSIDE the step's CUDA graph, on its own stream, behind the last weight-gradient launch of the bucket."""
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)

    def step(x, jt, vs, exchange=True):
        mu, wt = ops.joint_centers(jt, vs, (W // 4, H // 4), (W, H), 1)          # on-device Gaussian targets (A7)
        tgt = ops.gaussian_target(mu, wt, (W // 4, H // 4), 1)
        return eng.train_step(x, tgt, wt, lr, world_size=world, all_reduce=reduce_fn if (world > 1 and exchange) else None)

    for _ in range(warmup):
        loss = step(x_dev, j_dev, v_dev)
    ctx.barrier()
    ops.check_err_word(device)
    loss0 = float(loss)
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    if world > 1:
        for _ in range(10):                            # the step holds a collective: every rank runs the same count
            step(x_dev, j_dev, v_dev)                  # untimed: same load while nvidia-smi starts up
        torch.cuda.synchronize(device)
    else:
        t_wait = time.perf_counter()
        while sampler is not None and not sampler.started() and time.perf_counter() - t_wait < 3.0:
            step(x_dev, j_dev, v_dev)
            torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(steps):
        loss = step(x_dev, j_dev, v_dev)
    e1.record()
    ctx.barrier()
    clocks = sampler.stop() if sampler else None
    ms_step = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
    value = world * B / (ms_step * 1e-3)
    ops.check_err_word(device)
    loss1 = float(loss)
    # ---- end to end: pinned host images + joints -> H2D every step (on a side stream, one batch ahead), step, loss copied
    #      back every step and read by the host one step late (hgb200/prefetch.py)
    reader = LaggedScalar(device)                                  # staging buffers are set up once, before the clock
    feed = DevicePrefetcher(((host_x[i & 1], host_j, host_v) for i in range(steps)), device)
    feed.preallocate((host_x[0], host_j, host_v))
    ctx.barrier()
    t0 = time.perf_counter()
    for xd, jd, vd in feed:
        reader.push(step(xd, jd, vd))
    reader.flush()
    torch.cuda.synchronize(device)
    e2e_value = world * B * steps / ctx.max_over_ranks(time.perf_counter() - t0)
    ops.check_err_word(device)
    # ---- the exchange by itself, and what of it the step still sees (N > 1): (a) the whole flat buffer all-reduced
    #      eagerly, alone on the GPU, CUDA events; (b) the same K steps WITHOUT the collective -- the difference to ms_step
    #      is the exposed part.  Last, because after (b) the ranks' weights differ.
    allreduce_ms = allreduce_exposed_ms = 0.0
    ms_local = ms_step
    if world > 1:
        flat = eng.store.G[:eng.store.count]
        reduce_fn(flat)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        a.record()
        for _ in range(5):
            reduce_fn(flat)
        b.record()
        ctx.barrier()
        allreduce_ms = ctx.max_over_ranks(a.elapsed_time(b)) / 5
        for _ in range(2):
            step(x_dev, j_dev, v_dev, exchange=False)
        ctx.barrier()
        a.record()
        for _ in range(steps):
            step(x_dev, j_dev, v_dev, exchange=False)
        b.record()
        ctx.barrier()
        ms_local = ctx.max_over_ranks(a.elapsed_time(b)) / steps
        allreduce_exposed_ms = ms_step - ms_local
    # ---- roofline of the dominant kernel class
    hbm_peak, tf_burst, tf_sustained, peak_src = load_peaks()
    plan = eng.plans[(B, H, W)]
    per = plan.profile(iters=2)
    classes = class_table(per, plan.meta)
    total_ms = sum(per)
    # launch DAG (hgb200/dag.py): the longest dependency chain, priced with the per-launch eager times
    d, stream_of, waits = plan.schedule("step")
    fin = [0.0] * d.n
    for i in range(d.n):
        fin[i] = per[i] + max((fin[q] for q in d.preds[i]), default=0.0)
    dag_info = {"streams": _tr.STREAMS, "launch_closures": d.n, "edges": sum(len(q) for q in d.preds),
                "cross_stream_edges": sum(len(q) for q in waits), "sum_of_launch_ms": total_ms, "critical_path_ms": max(fin)}
    crit = {}
    i = max(range(d.n), key=lambda j: fin[j])
    while True:
        c = crit.setdefault(plan.meta[i]["op"], [0, 0.0])
        c[0] += 1
        c[1] += per[i]
        if not d.preds[i]:
            break
        i = max(d.preds[i], key=lambda j: fin[j])
    dag_info["critical_path_launches"] = sum(c[0] for c in crit.values())
    # the dominant kernel of the TRAINING step is the class that weighs most on the critical chain of the launch DAG:
    # leaves such as the weight-gradient GEMMs are deliberately confined to a quarter of the SMs and run beside it
    top_name = max(crit.items(), key=lambda kv: kv[1][1])[0]
    top = classes[top_name]
    roofline = roofline_of(top, top["ms"] / top["n"], hbm_peak, tf_sustained)
    roofline.update(traffic=ncu_traffic("train", top_name, B), kernel=top_name, launches_per_step=top["n"],
                    share_of_step=top["ms"] / total_ms, share_of_critical_chain=crit[top_name][1] / max(fin),
                    peak_source=f"{peak_src} (sustained bf16 / copy bandwidth, MEASURED_PEAKS.json)",
                    note="launch-inclusive eager timing; the step is bound by its chain of dependent launches, not by one "
                         "kernel (DESIGN.md section 5)")
    if args.train_breakdown and rank == 0:
        top_crit = sorted(crit.items(), key=lambda kv: -kv[1][1])[:40]
        write_breakdown(args.train_breakdown,
                        [f"# train step, per-kernel-class device time, eager replay with CUDA events, batch {B}; total "
                         f"{total_ms:.3f} ms; graph step {ms_step:.3f} ms", f"# launch DAG: {json.dumps(dag_info)}",
                         "# critical chain by class (launches, ms): " + "; ".join(f"{k} {v[0]} {v[1]:.3f}" for k, v in top_crit)],
                        classes, total_ms)
    launches = plan.num_kernel_launches
    eng.release_graphs()                  # graphs holding NCCL nodes must go before the process group does (Ctx.close)
    flops_per_step = TRAIN_GFLOP_PER_IMAGE * 1e9 * B
    tf = flops_per_step / (ms_step * 1e-3) / 1e12
    nbytes_grad = eng.store.count * 4
    return {
        "metric": TRAIN_METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": train_config(world, B),
        "loss_first_last": [loss0, loss1],
        "allreduce_ms": allreduce_ms, "allreduce_bytes": nbytes_grad if world > 1 else 0,
        "allreduce": {"overlapped_in_graph": bool(_tr.OVERLAP_ALLREDUCE) and world > 1,
                      "buckets": len(eng.grad_buckets()) if world > 1 else 0,
                      "alone_ms": allreduce_ms, "exposed_ms": allreduce_exposed_ms, "ms_per_step_without_exchange": ms_local,
                      "note": "alone_ms: the whole flat fp32 gradient buffer all-reduced eagerly with nothing else on the "
                              "GPU; exposed_ms: ms_per_step minus the same steps run without the collective"},
        "tensor_tflops": tf, "tensor_frac_of_measured_peak": tf / tf_sustained,
        "tensor_frac": {"of_sustained": tf / tf_sustained, "of_burst": tf / tf_burst, "of_nominal_2250": tf / 2250.0},
        "roofline": roofline, "launch_dag": dag_info,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * H * W * 4 + 2 * B * J * 3 * 8,
                "d2h_bytes_per_step": 4},
        "gpu_launches": (launches + 3) * steps,
        "clocks": clocks,
    }


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="inference: images per GPU per step")
    ap.add_argument("--train-batch", type=int, default=32, help="training: images per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--breakdown", default="", help="write the inference per-kernel-class time breakdown to this file")
    ap.add_argument("--train-breakdown", default="", help="write the training per-kernel-class time breakdown to this file")
    ap.add_argument("--workload", default="all", choices=["all", "infer", "train"],
                    help="all = C2 flip-test inference (headline) + C3 training step ('train' sub-record); or one of them")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    ctx = Ctx()
    infer = bench_infer(ctx, args) if args.workload in ("all", "infer") else None
    train = bench_train(ctx, args) if args.workload in ("all", "train") else None
    if ctx.rank != 0:
        ctx.close()
        return
    # ---------------- baselines (rank 0, N=1 only): bounded samples of the same workloads ----------------
    if ctx.world == 1 and not args.no_cpu_baseline:
        if infer is not None:
            n_img = 4
            rate, ms, cores, kind = cpu_reference_rate(n_img, steps=3, warmup=1)
            infer["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": cores, "kind": kind,
                                     "sample": f"{n_img} images/step, 3 steps: the reference's fp32 CPU flip-test forward + "
                                               f"flip average + get_final_preds_v1"}
        if train is not None:
            rate, ms, cores, kind = cpu_train_rate(2, 2)
            train["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": cores, "kind": kind,
                                     "sample": "2 images/step, 2 steps: the reference's fp32 CPU train step (autograd + RMSprop)"}
    else:
        for r in (infer, train):
            if r is not None:
                r["cpu_baseline"] = None
    gb = None
    if ctx.world == 1 and not args.no_gpu_baseline:
        gb = gpu_baseline(ctx.device, infer is not None, train is not None)
    if infer is not None:
        line = infer
        if train is not None:
            line["train"] = train
            line["gpu_launches"] += train["gpu_launches"]
            line["config"]["workload"] = WORKLOAD + BOTH + TRAIN_WORKLOAD + "  ['train' sub-record]"
    else:
        line = train
    line["gpu_baseline"] = gb
    print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
