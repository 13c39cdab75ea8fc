#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 stacked-hourglass hot path.

Workload (BASELINE.json configs[1]): 8-stack hourglass, MPII 16 joints, 256x256 input, batch 128 per GPU,
flip-test inference (two forwards per image), flip-average, arg-max + quarter-pixel + affine decode.
Metric: images/sec (original images, whole job over all GPUs).  One "step" = one batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our sm_100a path
  python bench.py --impl reference ...                           the reference algorithm on the host CPUs
                                                                (oracle port: /root/reference does not travel)

  python bench.py --workload train ...                            C3: 8-stack training step, batch 32 per GPU, JointsMSE
                                                                over all stacks, RMSprop; N>1 = data parallel with one
                                                                NCCL all-reduce of the flat gradient buffer per step

N>1 is launched by torchrun (one rank per GPU); inference shards by batch with no collective, so the only
torch.distributed traffic is the timing barrier / max-over-ranks.
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
METRIC = "images/sec 8-stack HG 256x256 flip-test inference"
WORKLOAD = "C2: 8-stack hourglass MPII 16-joint inference with flip test, batch 128, bf16"


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
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/r1_ncu_traffic.json), or None when no capture matches this kernel class and batch."""
    try:
        with open(os.path.join(REPO, "profiles", "r1_ncu_traffic.json")) as f:
            ent = json.load(f).get(workload, {}).get(f"{kernel_class}@{images}")
        return ent["dram_bytes"] if ent else None
    except (OSError, ValueError, KeyError):
        return None


def build_model(device, seed=0):
    """Random-init weights of the named architecture (torch default init, as the reference's constructor
    does) + randomised BN statistics so that folding is non-trivial."""
    import torch
    from src.models import hg
    torch.manual_seed(seed)
    model = hg(num_stacks=8, num_blocks=1, num_classes=16, mobile=False, skip_mode="sum", out_res=64)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.weight.copy_(0.75 + 0.5 * torch.rand(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))
    return model.to(device).eval()


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_reference_rate(n_images: int, steps: int, warmup: int):
    """The reference's algorithm (oracle port, fp32 NCHW torch on the host cores): flip-test forward of the
    8-stack network + flip-average + get_final_preds_v1 decode for `n_images` per step."""
    import numpy as np
    import torch
    from oracle.hourglass_oracle import make_state_dict, hg_forward
    from oracle import decode_oracle as D
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict(num_stacks=8, num_blocks=1, num_classes=16, seed=0)
    x = torch.randn(n_images, 3, 256, 256, generator=torch.Generator().manual_seed(2))
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
    return n_images / dt, dt * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img = 4
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    rate, ms, cores = cpu_reference_rate(n_img, steps, warmup)
    sample = f"{n_img} images/step (of the 128-image batch), flip-test forward + decode, fp32, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ training workload (C3)
TRAIN_GFLOP_PER_IMAGE = 3 * FWD_GFLOP_PER_IMAGE      # fwd + dgrad + wgrad
TRAIN_METRIC = "images/sec 8-stack HG 256x256 train"
TRAIN_WORKLOAD = "C3: 8-stack hourglass MPII training, intermediate-supervision JointsMSELoss, RMSprop, batch 32/GPU, bf16"


def cpu_train_rate(n_images: int, steps: int):
    """The reference's training step (oracle port: fp32 torch autograd on the host cores + RMSprop) on a bounded sample."""
    import torch
    from oracle.hourglass_oracle import make_state_dict
    from oracle import train_oracle as T
    from oracle.make_golden_inputs import train_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict(num_stacks=8, num_blocks=1, num_classes=16, seed=0)
    batches = train_inputs(1, n_images, 16, 256, 256, steps + 1)
    state = {}
    T.train_steps(sd, batches[:1], 2.5e-4)
    t0 = time.perf_counter()
    for b in batches[1:]:
        _, _, grads = T.forward_backward(sd, *b)
        T.rmsprop_update(sd, grads, state, 2.5e-4)
    dt = (time.perf_counter() - t0) / steps
    return n_images / dt, dt * 1e3, cores


def run_train(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        n_img, steps = 2, max(1, min(args.steps, 3))
        rate, ms, cores = cpu_train_rate(n_img, steps)
        sample = f"{n_img} images/step (of the 32-image batch), 8-stack train step fwd+bwd+RMSprop, fp32, {cores} threads"
        print(json.dumps({"impl": "reference", "metric": TRAIN_METRIC, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": TRAIN_WORKLOAD, "sample": sample},
                          "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    from hgb200 import lib, ops
    from hgb200.train import train_engine
    from src.models import hg
    lib.check(lib.hg_check_device(), "hg_check_device")
    steps, warmup = args.steps, max(3, args.warmup)
    B = args.batch if args.batch != 128 else 32
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
    reduce_fn = (lambda flat: dist.all_reduce(flat, op=dist.ReduceOp.SUM)) if world > 1 else None

    def step(x, jt, vs):
        mu, wt = ops.joint_centers(jt, vs, (W // 4, H // 4), (W, H), 1)          # on-device Gaussian targets (A7)
        tgt = ops.gaussian_target(mu, wt, (W // 4, H // 4), 1)
        return eng.train_step(x, tgt, wt, lr, world_size=world, all_reduce=reduce_fn)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    for _ in range(warmup):
        loss = step(x_dev, j_dev, v_dev)
    barrier()
    ops.check_err_word(device)
    loss0 = float(loss)
    sampler = ClockSampler(local_rank) if rank == 0 else None
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
    barrier()
    e0.record()
    for _ in range(steps):
        loss = step(x_dev, j_dev, v_dev)
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    value = world * B / (ms_step * 1e-3)
    ops.check_err_word(device)
    loss1 = float(loss)
    # ---- end to end: pinned host images + joints -> H2D every step (on a side stream, one batch ahead), step, loss copied
    #      back every step and read by the host one step late (hgb200/prefetch.py)
    from hgb200.prefetch import DevicePrefetcher, LaggedScalar
    reader = LaggedScalar(device)                                  # staging buffers are set up once, before the clock
    feed = DevicePrefetcher(((host_x[i & 1], host_j, host_v) for i in range(steps)), device)
    feed.preallocate((host_x[0], host_j, host_v))
    barrier()
    t0 = time.perf_counter()
    for xd, jd, vd in feed:
        lv = reader.push(step(xd, jd, vd))
    lv = reader.flush()
    torch.cuda.synchronize(device)
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * steps / float(te.item())
    # ---- roofline of the dominant kernel class
    hbm_peak, tf_burst, tf_sustained, peak_src = load_peaks()
    plan = eng.plans[(B, H, W)]
    per = plan.profile(iters=2)
    classes = {}
    for ms, meta in zip(per, plan.meta):
        c = classes.setdefault(meta["op"], dict(ms=0.0, n=0, flops=meta["flops"], bytes=meta["bytes"], kind=meta["kind"]))
        c["ms"] += ms
        c["n"] += 1
    total_ms = sum(per)
    # launch DAG (hgb200/dag.py): the longest dependency chain, priced with the per-launch eager times
    import hgb200.train as _tr
    d, stream_of, waits = plan.schedule("step")
    fin = [0.0] * d.n
    for i in range(d.n):
        fin[i] = per[i] + max((fin[q] for q in d.preds[i]), default=0.0)
    dag_info = {"streams": _tr.STREAMS, "launch_closures": d.n, "edges": sum(len(q) for q in d.preds),
                "cross_stream_edges": sum(len(q) for q in waits), "sum_of_launch_ms": total_ms, "critical_path_ms": max(fin)}
    # walk the critical chain back from its end: which kernel classes it is made of
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
    avg_ms = top["ms"] / top["n"]
    if top["kind"] == "conv" and top["flops"] / max(top["bytes"], 1) > tf_sustained * 1e12 / (hbm_peak * 1e9):
        roofline = {"bound": "tensor", "achieved": top["flops"] / (avg_ms * 1e-3) / 1e12, "peak": tf_sustained, "unit": "TFLOP/s"}
    else:
        roofline = {"bound": "hbm", "achieved": top["bytes"] / (avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"}
    roofline.update(frac=roofline["achieved"] / roofline["peak"], traffic=None, kernel=top_name, launches_per_step=top["n"],
                    share_of_step=top["ms"] / total_ms, share_of_critical_chain=crit[top_name][1] / max(fin),
                    peak_source=f"{peak_src} (sustained bf16 / copy bandwidth, MEASURED_PEAKS.json)",
                    note="launch-inclusive eager timing of a 15-60 us kernel; the step is bound by the chain of ~1700 "
                         "dependent launches, not by one kernel (DESIGN.md section 5)")
    if args.breakdown and rank == 0:
        with open(args.breakdown, "w") as f:
            f.write(f"# train step, per-kernel-class device time, eager replay with CUDA events, batch {B}; total {total_ms:.3f} ms; "
                    f"graph step {ms_step:.3f} ms\n")
            f.write(f"# launch DAG: {json.dumps(dag_info)}\n")
            top_crit = sorted(crit.items(), key=lambda kv: -kv[1][1])[:40]
            f.write("# critical chain by class (launches, ms): " + "; ".join(f"{k} {v[0]} {v[1]:.3f}" for k, v in top_crit) + "\n")
            f.write("class,launches,total_ms,avg_ms,share,TFLOP/s,GB/s\n")
            for name, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"]):
                a = c["ms"] / c["n"]
                f.write(f"{name},{c['n']},{c['ms']:.4f},{a:.4f},{c['ms']/total_ms:.4f},"
                        f"{c['flops']/(a*1e-3)/1e12:.1f},{c['bytes']/(a*1e-3)/1e9:.0f}\n")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, ms, cores = cpu_train_rate(2, 2)
        cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "2 images/step, 2 steps: oracle port of the reference train step (fp32 torch CPU autograd + RMSprop)"}
    flops_per_step = TRAIN_GFLOP_PER_IMAGE * 1e9 * B
    line = {
        "metric": TRAIN_METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": TRAIN_WORKLOAD, "images_per_gpu_per_step": B,
                   "parallelism": f"data parallel x{world}, one NCCL all-reduce of the flat fp32 gradient buffer per step",
                   "l2": "working set (>9 GB of saved activations per step) far exceeds the 126 MB L2; no flush needed",
                   "weights": "random init (torch default)", "loss_first_last": [loss0, loss1]},
        "tensor_tflops": flops_per_step / (ms_step * 1e-3) / 1e12,
        "tensor_frac_of_measured_peak": flops_per_step / (ms_step * 1e-3) / 1e12 / tf_sustained,
        "roofline": roofline, "cpu_baseline": cpu, "launch_dag": dag_info,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * H * W * 4 + 2 * B * J * 3 * 8,
                "d2h_bytes_per_step": 4},
        "gpu_launches": (plan.num_kernel_launches + 3) * steps,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="images per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", default="", help="write the per-kernel-class time breakdown to this file")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer = C2 flip-test inference (headline); train = C3 training step")
    args = ap.parse_args()
    if args.workload == "train":
        run_train(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from hgb200 import lib, ops
    from hgb200.infer import FlipTestPipeline
    lib.check(lib.hg_check_device(), "hg_check_device")

    steps, warmup = args.steps, max(3, args.warmup)
    B = args.batch
    model = build_model(device)
    engine = model.engine(device)
    pipe = FlipTestPipeline(engine, B, 256, 256)
    pipe.set_affine(np.tile([[128.0, 128.0]], (B, 1)), np.tile([[1.28, 1.28]], (B, 1)))
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    host = [torch.randn(B, 3, 256, 256, generator=g).pin_memory() for _ in range(2)]
    x_dev = host[0].to(device)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    # ---------------- device-resident throughput (inputs already in HBM) ----------------
    for _ in range(warmup):
        pipe.infer_device(x_dev)
    barrier()
    ops.check_err_word(device)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_wait = time.perf_counter()
    while sampler is not None and not sampler.started() and time.perf_counter() - t_wait < 3.0:
        pipe.infer_device(x_dev)                       # untimed: same load while nvidia-smi starts up
        torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        pipe.infer_device(x_dev)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    value = world * B / (ms_step * 1e-3)
    ops.check_err_word(device)

    # ---------------- end to end: pinned host input -> H2D -> graph -> D2H coordinates ----------------
    def host_batches(k):
        for i in range(k):
            yield host[i & 1]

    for _ in pipe.infer_host(host_batches(2)):
        pass
    barrier()
    t0 = time.perf_counter()
    n_out = 0
    for coords in pipe.infer_host(host_batches(steps)):
        n_out += coords.shape[0]
    torch.cuda.synchronize(device)
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * steps / float(te.item())
    assert n_out == B * steps

    # ---------------- roofline of the dominant kernel (CUDA events around every launch, eager replay) ----------
    hbm_peak, tf_burst, tf_sustained, peak_src = load_peaks()
    per_launch = pipe.plan.profile(iters=2)
    classes = {}
    for ms, meta in zip(per_launch, pipe.plan.meta):
        c = classes.setdefault(meta["op"], dict(ms=0.0, n=0, flops=meta["flops"], bytes=meta["bytes"], kind=meta["kind"]))
        c["ms"] += ms
        c["n"] += 1
    total_ms = sum(per_launch)
    top_name, top = max(classes.items(), key=lambda kv: kv[1]["ms"])
    avg_ms = top["ms"] / top["n"]
    if top["kind"] == "conv" and top["flops"] / max(top["bytes"], 1) > tf_sustained * 1e12 / (hbm_peak * 1e9):
        roofline = {"bound": "tensor", "achieved": top["flops"] / (avg_ms * 1e-3) / 1e12, "peak": tf_sustained,
                    "unit": "TFLOP/s"}
    else:
        roofline = {"bound": "hbm", "achieved": top["bytes"] / (avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"}
    roofline["frac"] = roofline["achieved"] / roofline["peak"]
    roofline["traffic"] = ncu_traffic("infer", top_name, B)
    roofline["kernel"] = top_name
    roofline["launches_per_step"] = top["n"]
    roofline["share_of_step"] = top["ms"] / total_ms
    roofline["peak_source"] = f"{peak_src} (sustained bf16 / copy bandwidth, MEASURED_PEAKS.json)"
    if args.breakdown and rank == 0:
        with open(args.breakdown, "w") as f:
            f.write(f"# per-kernel-class device time, eager replay with CUDA events, batch {B} (x2 flip) ; total {total_ms:.3f} ms\n")
            f.write("class,launches,total_ms,avg_ms,share,TFLOP/s,GB/s\n")
            for name, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"]):
                a = c["ms"] / c["n"]
                f.write(f"{name},{c['n']},{c['ms']:.4f},{a:.4f},{c['ms']/total_ms:.4f},"
                        f"{c['flops']/(a*1e-3)/1e12:.1f},{c['bytes']/(a*1e-3)/1e9:.0f}\n")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_img = 4
        rate, ms, cores = cpu_reference_rate(n_img, steps=3, warmup=1)
        cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{n_img} images/step, 3 steps: oracle port of the reference (fp32 torch CPU) flip-test forward + decode"}

    flops_per_step = 2 * FWD_GFLOP_PER_IMAGE * 1e9 * B     # two forwards per image
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_gpu_per_step": B, "forwards_per_image": 2,
                   "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": "working set (>3 GB of activations per step) far exceeds the 126 MB L2; no flush needed",
                   "note": "value (device-resident inputs) and e2e (pinned host inputs, H2D double-buffered behind compute) are "
                           "timed in separate back-to-back regions; under sw_power_cap the SM clock drifts a few percent "
                           "between them, so e2e can land on either side of value",
                   "weights": "random init (torch default), randomised BN statistics"},
        "tensor_tflops": flops_per_step / (ms_step * 1e-3) / 1e12,
        "tensor_frac_of_measured_peak": flops_per_step / (ms_step * 1e-3) / 1e12 / tf_sustained,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * 256 * 256 * 4,
                "d2h_bytes_per_step": B * 16 * 2 * 8},
        "gpu_launches": pipe.launches_per_batch * steps,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
