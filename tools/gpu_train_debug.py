"""Verbose sweep of the training kernels and of one whole training step against torch / the CPU oracle.
Keeps going after a failure so one gpurun call yields the full picture.  Not a pytest file."""
import os
import sys
import time
import traceback

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "hourglass-pose-estimation_b200"),
                os.path.join(os.path.dirname(HERE), "tests")]
from hgb200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def r16(x):
    return x.to(torch.bfloat16).float()


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def check(name, fn):
    try:
        msg = fn()
        print(f"[{name}] {msg}", flush=True)
    except Exception:  # noqa: BLE001
        print(f"[{name}] EXC\n{traceback.format_exc()}", flush=True)


def t_bn(n, h, w, c, halo):
    g = torch.Generator().manual_seed(c + h)
    x = (torch.randn(n, h, w, c, generator=g) * 1.5 + 0.3).to(torch.bfloat16).cuda()
    gamma = (0.5 + torch.rand(c, generator=g)).cuda()
    beta = (0.2 * torch.randn(c, generator=g)).cuda()
    rm, rv = torch.zeros(c).cuda(), torch.ones(c).cuda()
    nbt = torch.zeros((), dtype=torch.int64).cuda()
    sums = torch.zeros(2 * c).cuda()
    saved = torch.zeros(4 * c).cuda()
    ops.colstats(x, sums[:c], sums[c:])
    out = ops.halo_padded_buffer(n, h, w, c, "cuda") if halo else torch.empty_like(x)
    ops.bn_train_fwd(x, sums, gamma, beta, rm, rv, nbt, saved, out, halo=halo)
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    rm2, rv2 = torch.zeros(c).cuda(), torch.ones(c).cuda()
    gp, bp = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.relu(F.batch_norm(xf, rm2, rv2, gp, bp, True, 0.1, 1e-5))
    o = ops.halo_interior(out, n, h, w, c) if halo else out
    e_f = rel(o.float().permute(0, 3, 1, 2), ref.detach())
    e_rm, e_rv = rel(rm, rm2), rel(rv, rv2)
    # backward
    dz = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    add1 = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    bs = torch.zeros(2 * c).cuda()
    ops.bn_bwd_reduce(dz, x, saved, bs)
    dgam, dbet = torch.zeros(c).cuda(), torch.zeros(c).cuda()
    dx = ops.halo_padded_buffer(n, h, w, c, "cuda") if halo else torch.empty_like(x)
    ops.bn_bwd_apply(dz, x, saved, bs, dx, add1=None if halo else add1, dgamma=dgam, dbeta=dbet, halo=halo)
    ref.backward(dz.float().permute(0, 3, 1, 2))
    dxr = xf.grad + (0 if halo else add1.float().permute(0, 3, 1, 2))
    d = ops.halo_interior(dx, n, h, w, c) if halo else dx
    e_dx = rel(d.float().permute(0, 3, 1, 2), dxr)
    e_g, e_b = rel(dgam, gp.grad), rel(dbet, bp.grad)
    pads_ok = True
    if halo:
        full = dx.clone()
        ops.halo_interior(full, n, h, w, c).zero_()
        pads_ok = bool((full == 0).all()) and int(nbt) == 1
    ok = max(e_f, e_dx) < 1.5e-2 and max(e_rm, e_rv, e_g, e_b) < 2e-3 and pads_ok
    return f"{n}x{h}x{w}x{c} halo={halo}: fwd {e_f:.2e} rm {e_rm:.1e} rv {e_rv:.1e} dx {e_dx:.2e} dgamma {e_g:.1e} dbeta {e_b:.1e} pads {pads_ok} {'OK' if ok else 'FAIL'}"


def t_pool(n, h, w, c):
    g = torch.Generator().manual_seed(h)
    x = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    dp = torch.randn(n, h // 2, w // 2, c, generator=g).to(torch.bfloat16).cuda()
    base = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.max_pool2d(xf, 2, 2).backward(dp.float().permute(0, 3, 1, 2))
    dx = torch.empty_like(x)
    ops.maxpool2x2_bwd(x, dp, dx, False)
    e0 = rel(dx.float().permute(0, 3, 1, 2), xf.grad)
    dx2 = base.clone()
    ops.maxpool2x2_bwd(x, dp, dx2, True)
    e1 = rel(dx2.float().permute(0, 3, 1, 2), r16(xf.grad + base.float().permute(0, 3, 1, 2)))
    dy = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    lo = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device="cuda")
    ops.sumpool2x2(dy, lo)
    ref = F.avg_pool2d(dy.float().permute(0, 3, 1, 2), 2) * 4
    e2 = rel(lo.float().permute(0, 3, 1, 2), ref)
    a = base.clone()
    ops.add_inplace(a, dy)
    e3 = rel(a.float(), r16(base.float() + dy.float()))
    ok = e0 == 0 and e1 < 1e-6 and e2 < 8e-3 and e3 < 1e-6
    return f"pool bwd {e0:.1e} acc {e1:.1e} sumpool {e2:.1e} add {e3:.1e} {'OK' if ok else 'FAIL'}"


def t_pack_rms():
    g = torch.Generator().manual_seed(3)
    co, taps, ci = 24, 9, 64
    src = torch.randn(co * taps * ci, generator=g).cuda()
    fwd = torch.zeros(32, taps * ci + 64, dtype=torch.bfloat16, device="cuda")
    dg = torch.zeros(ci, 256, dtype=torch.bfloat16, device="cuda")
    tab = ops.make_pack_table([dict(src=src, dst_fwd=fwd, dst_dgrad=dg, co=co, taps=taps, ci=ci, fwd_ld=taps * ci + 64,
                                    fwd_col0=64, dgrad_ld=256)], "cuda")
    ops.pack_weights(tab, 1)
    w = src.view(co, taps, ci)
    e0 = rel(fwd[:co, 64:].float(), r16(w.reshape(co, -1)))
    ref_d = r16(w.flip(1).permute(2, 1, 0).reshape(ci, taps * co))
    e1 = rel(dg[:, :taps * co].float(), ref_d)
    p = torch.randn(1000, generator=g).cuda()
    gr = torch.randn(1000, generator=g).cuda()
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.RMSprop([pt], lr=1e-3)
    v = torch.zeros(1000).cuda()
    for _ in range(3):
        pt.grad = gr.clone()
        opt.step()
        ops.rmsprop_step(p, gr, v, 1e-3)
    e2 = rel(p, pt.detach())
    A, Bm = torch.randn(7, 5, generator=g).cuda(), torch.randn(5, 9, generator=g).cuda()
    D = torch.randn(7, 9, generator=g).cuda()
    Cm = torch.ones(7, 9).cuda()
    ops.small_gemm(Cm, A, Bm, D, 7, 9, 5, 5, 1, 9, 1, 9, 1, beta=1.0)
    e3 = rel(Cm, 1 + D + A @ Bm)
    ok = e0 == 0 and e1 == 0 and e2 < 1e-5 and e3 < 1e-5
    return f"pack fwd {e0:.1e} dgrad {e1:.1e} rmsprop {e2:.1e} small_gemm {e3:.1e} {'OK' if ok else 'FAIL'}"


def load_emulation():
    """A second instance of hgb200.train whose `ops` is the CPU torch emulation (tests/fake_ops.py)."""
    import importlib.util
    import fake_ops
    path = os.path.join(os.path.dirname(HERE), "hourglass-pose-estimation_b200", "hgb200", "train.py")
    spec = importlib.util.spec_from_file_location("hgb200.train_emulation", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.ops = fake_ops
    return mod


def grad_rows(grads, ref_grads):
    rows = []
    for k, gref in ref_grads.items():
        gm, gr = grads[k].reshape(-1).double(), gref.reshape(-1).double()
        nr = float(gr.norm())
        rows.append((float((gm * gr).sum() / (gm.norm() * gr.norm() + 1e-300)), float((gm - gr).norm() / (nr + 1e-300)), nr, k))
    return rows


def summarize(tag, rows):
    gmax = max(w[2] for w in rows)
    sig = sorted(w for w in rows if w[2] > 1e-4 * gmax)
    return (f"{tag}: {len(sig)} significant grads, min cos {sig[0][0]:.4f}, median relL2 "
            f"{np.median([w[1] for w in sig]):.3e}, max relL2 {max(w[1] for w in sig):.3e}"), sig


def t_step(S, J, B, H, W, steps=2, lr=2.5e-4, use_graph=True, autograd=False, emulate=True):
    from src.models import hg
    from src.loss import MSELoss
    from oracle.hourglass_oracle import make_state_dict
    from oracle import train_oracle as T
    from oracle.make_golden_inputs import train_inputs
    from test_train_plan_cpu import autocast_yardstick
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    sd = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0)
    model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum")
    model.load_state_dict(sd)
    model = model.cuda().train()
    batches = train_inputs(1, B, J, H, W, steps)
    from hgb200.train import train_engine
    eng = train_engine(model)
    emu = emu_model = None
    if emulate:
        em = load_emulation()
        emu_model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum")
        emu_model.load_state_dict(sd)
        emu_model.train()
        emu = em.TrainEngine(emu_model, "cpu")
    sd_ref = {k: v.clone() for k, v in sd.items()}
    state = {}
    lines = []
    opt = torch.optim.RMSprop(model.parameters(), lr=lr) if autograd else None
    crit = MSELoss(use_target_weight=True)
    for step, (x, tg, tw) in enumerate(batches):
        t0 = time.time()
        ac_loss, ac_outs, ac_grads = autocast_yardstick(sd_ref, x, tg, tw)
        ref_loss, ref_outs, ref_grads = T.forward_backward(sd_ref, x, tg, tw)
        t_ref = time.time() - t0
        if autograd:
            outs = model(x.cuda())
            loss_t = crit(outs, tg.cuda(), tw.cuda())
            opt.zero_grad()
            loss_t.backward()
            loss = float(loss_t)
        else:
            plan = eng.plan_for(B, H, W)
            plan.input.copy_(x)
            plan.target.copy_(tg)
            plan.target_weight.copy_(tw.reshape(B, J))
            plan.run("step", use_graph)
            torch.cuda.synchronize()
            ops.check_err_word()
            loss = float(plan.loss)
            outs = plan.outputs
        mine = {k: model.state_dict(keep_vars=True)[k].grad.detach().cpu().contiguous() for k in ref_grads}
        hm_err = max(rel(o.cpu(), r) for o, r in zip(outs, ref_outs))
        hm_ac = max(rel(o.float(), r) for o, r in zip(ac_outs, ref_outs))
        lines.append(f"  step {step}: loss {loss:.6e} ref {ref_loss:.6e} (rel {abs(loss - ref_loss) / ref_loss:.2e}; autocast "
                     f"{abs(ac_loss - ref_loss) / ref_loss:.2e}) heatmap err vs fp32 {hm_err:.3e} (autocast {hm_ac:.3e}) oracle {t_ref:.1f}s")
        msg, sig = summarize("      vs fp32 oracle", grad_rows(mine, ref_grads))
        lines.append(msg)
        lines.append(summarize("      stock autocast vs fp32", grad_rows(ac_grads, ref_grads))[0])
        if emu is not None:
            t0 = time.time()
            eplan = emu.plan_for(B, H, W)
            eplan.input.copy_(x)
            eplan.target.copy_(tg)
            eplan.target_weight.copy_(tw.reshape(B, J))
            eplan.run("step", False)
            egr = {k: emu_model.state_dict(keep_vars=True)[k].grad.detach().contiguous().clone() for k in ref_grads}
            hm_e = max(rel(o.cpu(), r) for o, r in zip(outs, eplan.outputs))
            msg, sige = summarize(f"      vs bf16 emulation (loss rel {abs(loss - float(eplan.loss)) / float(eplan.loss):.2e}, "
                                  f"heatmap {hm_e:.3e}, {time.time() - t0:.1f}s)", grad_rows(mine, egr))
            lines.append(msg)
            for cos, l2, nr, k in sige[:5]:
                lines.append(f"          {k}: cos {cos:.4f} relL2 {l2:.3e} |g| {nr:.3e}")
        # optimizer
        T.rmsprop_update(sd_ref, ref_grads, state, lr)
        if autograd:
            opt.step()
        else:
            eng.rmsprop(lr)
        torch.cuda.synchronize()
        d = torch.cat([(model.state_dict()[k].detach().cpu() - sd_ref[k]).abs().reshape(-1) for k in ref_grads])
        bn_err = max(rel(model.state_dict()[k].cpu().float(), sd_ref[k].float()) for k in sd_ref if k.endswith("running_var"))
        lines.append(f"      after update: |dparam| max {float(d.max()):.3e} median {float(d.median()):.3e} (one step moves "
                     f"{lr / np.sqrt(1 - 0.99):.2e}); running_var rel err {bn_err:.2e}")
        model.load_state_dict(sd_ref)          # continue from the oracle's parameters
        if emu is not None:
            emu_model.load_state_dict(sd_ref)
        if autograd:
            opt = torch.optim.RMSprop(model.parameters(), lr=lr)
    return "\n" + "\n".join(lines)


def main():
    which = sys.argv[1:] or ["kernels", "step"]
    if "kernels" in which:
        for args in [(2, 16, 16, 64, False), (2, 16, 16, 128, True), (3, 8, 8, 256, False), (2, 64, 48, 128, True),
                     (5, 4, 4, 128, True), (1, 128, 128, 64, False), (32, 4, 4, 256, False)]:
            check("bn", lambda a=args: t_bn(*a))
        check("pool", lambda: t_pool(2, 16, 16, 256))
        check("pool", lambda: t_pool(3, 4, 6, 64))
        check("pack", t_pack_rms)
    if "step" in which:
        check("step S1 64x64 eager", lambda: t_step(1, 16, 2, 64, 64, steps=1, use_graph=False))
        check("step S2 128x128 graph", lambda: t_step(2, 16, 4, 128, 128, steps=2, use_graph=True))
        check("step S2 128x128 autograd", lambda: t_step(2, 16, 4, 128, 128, steps=2, autograd=True))
        check("step S2 j17 128x192", lambda: t_step(2, 17, 2, 128, 192, steps=1, use_graph=True))


if __name__ == "__main__":
    main()
