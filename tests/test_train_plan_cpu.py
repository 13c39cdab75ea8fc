"""Host logic of the training plan (hgb200/train.py) on CPU: the plan's launch sequence is executed through
tests/fake_ops.py (a plain-torch emulation of every libhgb200 entry point the plan drives, same bf16 rounding
points) and compared with the fp32 training oracle.  Checks gradient routing through the residual stream,
the hourglass skip connections and pools, buffer aliasing/recycling, weight packing (incl. the tap-flipped
dgrad weights), the merged remap convolution's parameter-space chain rule, BN running statistics and RMSprop."""
import numpy as np
import pytest
import torch

import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fake_ops  # noqa: E402
from oracle.hourglass_oracle import make_state_dict
from oracle import train_oracle as T
from oracle.make_golden_inputs import train_inputs


@pytest.fixture()
def cpu_train(monkeypatch):
    import hgb200.train as tr
    monkeypatch.setattr(tr, "ops", fake_ops)
    return tr


def _grad_report(model, ref_grads):
    rows = []
    sdk = model.state_dict(keep_vars=True)
    for k, gref in ref_grads.items():
        gm = sdk[k].grad.detach().contiguous().reshape(-1).double()
        gr = gref.reshape(-1).double()
        cos = float((gm * gr).sum() / (gm.norm() * gr.norm() + 1e-300))
        rows.append((cos, float((gm - gr).norm() / (gr.norm() + 1e-300)), float(gr.norm()), k))
    return rows


@pytest.mark.parametrize("S,J,B,H,W,mobile,skip", [(2, 16, 4, 128, 128, False, "sum"), (1, 17, 3, 64, 128, False, "sum"),
                                                   (2, 16, 2, 128, 128, True, "sum"), (1, 16, 4, 128, 128, False, "concat"),
                                                   (1, 14, 4, 128, 128, True, "concat")])
def test_plan_is_exact_with_fp32_storage(cpu_train, monkeypatch, S, J, B, H, W, mobile, skip):
    """With activations and GEMM weights stored in fp32 the plan IS the reference's step: every gradient must
    match the oracle to accumulation-order noise.  (bf16 storage is the product's numeric type; its noise is
    measured against a stock-PyTorch bf16 yardstick in tests/test_gpu_train.py.)"""
    from src.models import hg
    monkeypatch.setattr(cpu_train, "_ACT", torch.float32)
    monkeypatch.setattr(fake_ops, "BF", torch.float32)
    monkeypatch.setattr(cpu_train, "FUSED_STATS", skip == "concat")     # both ways of getting the BN statistics
    sd = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, mobile=mobile, skip_mode=skip, seed=0)
    model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=mobile, skip_mode=skip)
    model.load_state_dict(sd)
    model.train()
    eng = cpu_train.TrainEngine(model, "cpu")
    # parameters are now channels_last views of the flat buffer, values unchanged
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k
    lr = 2.5e-4
    sd_ref = {k: v.clone() for k, v in sd.items()}
    state = {}
    for step, (x, tg, tw) in enumerate(train_inputs(1, B, J, H, W, 2)):
        ref_loss, ref_outs, ref_grads = T.forward_backward(sd_ref, x, tg, tw)
        loss = eng.train_step(x, tg, tw, lr, use_graph=False)
        plan = eng.plans[(B, H, W)]
        assert abs(float(loss) - ref_loss) <= 1e-4 * ref_loss
        for o, r in zip(plan.outputs, ref_outs):
            assert float((o - r).abs().max()) <= 2e-3 * float(r.abs().max())
        rows = _grad_report(model, ref_grads)
        gmax = max(r[2] for r in rows)
        sig = [r for r in rows if r[2] > 1e-4 * gmax]
        assert len(sig) > 0.5 * len(rows)
        bad = [r for r in sig if r[0] < 0.999 or r[1] > 5e-2]
        assert not bad, sorted(bad)[:8]
        # conv biases feeding a train-mode BN have analytically zero gradient: the plan leaves exact zeros there
        for k in ref_grads:
            if k.endswith("conv1.bias") or k.endswith("conv2.bias"):
                assert float(model.state_dict(keep_vars=True)[k].grad.abs().max()) == 0.0
        T.rmsprop_update(sd_ref, ref_grads, state, lr)
        for k in sd_ref:
            if k.endswith("running_mean") or k.endswith("running_var"):
                np.testing.assert_allclose(model.state_dict()[k].numpy(), sd_ref[k].numpy(), rtol=1e-3, atol=1e-4)
        assert int(model.bn1.num_batches_tracked) == step + 1
        # RMSprop: same update rule.  Early steps move every weight by ~lr/sqrt(1-alpha) whatever the gradient's size,
        # so noise-level gradients may step either way: compare the bulk tightly and everything within the step size
        step_max = (step + 1) * 1.05 * lr / np.sqrt(1 - T.RMSPROP_ALPHA)
        diffs = []
        for k in ref_grads:
            d = (model.state_dict()[k] - sd_ref[k]).abs()
            assert float(d.max()) <= 2 * step_max, k
            diffs.append(d.reshape(-1))
        assert float(torch.cat(diffs).median()) <= 1e-5
        # continue from the oracle's parameters (also exercises load_state_dict through the flat-buffer views)
        model.load_state_dict(sd_ref)
        eng.store.V.zero_()
        for k, v in state.items():
            eng.store.view(eng.store.V, k).copy_(v)


def autocast_yardstick(sd, x, tg, tw):
    """The same step through stock PyTorch bf16 autocast (CPU): what 'standard bf16 training numerics' gives
    against fp32 on this input.  Returns (loss, outputs, grads)."""
    sd2 = {k: v.clone() for k, v in sd.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        return T.forward_backward(sd2, x, tg, tw)


def rel_l2_rows(grads, ref_grads):
    rows = []
    for k, gref in ref_grads.items():
        gm, gr = grads[k].reshape(-1).double(), gref.reshape(-1).double()
        rows.append((float((gm - gr).norm() / (gr.norm() + 1e-300)), float(gr.norm()), k))
    return rows


def test_plan_bf16_storage_is_no_noisier_than_stock_bf16_autocast(cpu_train):
    """Product numeric type (bf16 storage, fp32 accumulate) through the same plan.  Train-mode BatchNorm
    re-normalises every conv output, so bf16 rounding noise is amplified block after block on a randomly
    initialised network (stock PyTorch autocast shows heat maps ~15 % off fp32 on this input); the yardstick
    is therefore stock bf16 autocast itself: the plan must not be noisier than that against the fp32 oracle."""
    from src.models import hg
    S, J, B, H, W = 1, 16, 4, 128, 128
    sd = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0)
    model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum")
    model.load_state_dict(sd)
    model.train()
    eng = cpu_train.TrainEngine(model, "cpu")
    x, tg, tw = train_inputs(1, B, J, H, W, 1)[0]
    sd_ref = {k: v.clone() for k, v in sd.items()}
    ref_loss, ref_outs, ref_grads = T.forward_backward(sd_ref, x, tg, tw)
    ac_loss, ac_outs, ac_grads = autocast_yardstick(sd, x, tg, tw)
    loss = eng.train_step(x, tg, tw, 2.5e-4, use_graph=False)
    assert abs(float(loss) - ref_loss) <= 2e-2 * ref_loss
    plan = eng.plans[(B, H, W)]
    hm = float((plan.outputs[-1] - ref_outs[-1]).abs().max())
    hm_ac = float((ac_outs[-1].float() - ref_outs[-1]).abs().max())
    assert hm <= 1.25 * hm_ac
    mine = {k: model.state_dict(keep_vars=True)[k].grad.detach().contiguous() for k in ref_grads}
    gmax = max(float(g.norm()) for g in ref_grads.values())
    keep = [k for k, g in ref_grads.items() if float(g.norm()) > 1e-4 * gmax]
    m_mine = np.median([r[0] for r in rel_l2_rows(mine, ref_grads) if r[2] in keep])
    m_ac = np.median([r[0] for r in rel_l2_rows(ac_grads, ref_grads) if r[2] in keep])
    assert m_mine <= 1.15 * m_ac, (m_mine, m_ac)


def test_cpu_model_training_fails_loudly():
    from src.models import hg
    model = hg(num_stacks=1, num_blocks=1, num_classes=16, mobile=False, skip_mode="sum").train()
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 64, 64))


def test_autograd_dropin_forward_may_be_repeated_but_a_stale_backward_is_refused(cpu_train, monkeypatch):
    """model(x) in train mode under the reference's own loop (trainer.py:89-99).  The saved activations are the plan's static
    buffers: a forward whose backward never runs is fine (logging / evaluation in train mode), the backward of a forward
    that is no longer the latest of its shape must fail loudly instead of using another batch's activations."""
    from hgb200 import HgError
    from src.models import hg
    monkeypatch.setattr(cpu_train, "_ACT", torch.float32)
    monkeypatch.setattr(fake_ops, "BF", torch.float32)
    S, J, B, H, W = 1, 16, 3, 128, 128      # (at 64x64 the lowest level is 1x1: batch statistics over B values are degenerate)
    model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum")
    model.load_state_dict(make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0))
    model.train()
    model.use_cuda_graph = False
    (x, tg, tw), (x2, _, _) = train_inputs(1, B, J, H, W, 2)

    def crit(outs, target, weight):
        """src/loss/mse.py:14-44 in plain torch ops (the product's MSELoss is a CUDA kernel)."""
        total = 0.0
        for o in outs:
            b, j = o.shape[:2]
            p, g = o.reshape(b, j, -1), target.reshape(b, j, -1)
            per = sum(0.5 * torch.mean((p[:, k] * weight[:, k] - g[:, k] * weight[:, k]) ** 2) for k in range(j))
            total = total + per / j
        return total

    out_a = model(x)                     # never back-propagated
    out_b = model(x2)
    out_c = model(x)                     # the latest forward of this shape
    assert all(torch.equal(a, c) for a, c in zip(out_a, out_c)) and not torch.equal(out_b[-1], out_c[-1])
    ref_loss, _, ref_grads = T.forward_backward({k: v.detach().clone() for k, v in model.state_dict().items()}, x, tg, tw)
    loss = crit(out_c, tg, tw)
    loss.backward()
    assert abs(float(loss) - ref_loss) <= 1e-4 * ref_loss
    g = model.score[0].weight.grad
    assert float((g - ref_grads["score.0.weight"]).norm()) <= 1e-3 * float(ref_grads["score.0.weight"].norm())
    with pytest.raises(HgError):
        crit(out_b, tg, tw).backward()   # out_b's activations were overwritten by the forward that produced out_c
