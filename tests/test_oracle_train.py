"""The training oracle (oracle/train_oracle.py) against the live reference's two-step fixture
(tests/golden/train_s2_j16_128.npz, written by oracle/make_golden.py:train_golden)."""
import os

import numpy as np
import torch

from oracle.hourglass_oracle import make_state_dict
from oracle import train_oracle as T
from oracle.make_golden_inputs import train_inputs, TRAIN_STRIDE


def test_train_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_s2_j16_128.npz"))
    S, J, B, H, W, seed, steps = [int(v) for v in g["cfg"]]
    assert int(g["stride"]) == TRAIN_STRIDE
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    sd = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=seed)
    losses, grads, _ = T.train_steps(sd, train_inputs(seed + 1, B, J, H, W, steps), lr=float(g["lr"]))
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-5)
    keys = sorted(grads)
    gs = np.concatenate([grads[k].reshape(-1).numpy() for k in keys])[::TRAIN_STRIDE]
    scale = np.abs(g["grad_sample"]).max()
    assert np.abs(gs - g["grad_sample"]).max() <= 2e-4 * scale
    norms = np.array([float(grads[k].norm()) for k in keys])
    # conv biases feeding a train-mode BN have analytically zero gradient: only roundoff there
    big = g["grad_norms"] > 1e-6 * g["grad_norms"].max()
    np.testing.assert_allclose(norms[big], g["grad_norms"][big], rtol=5e-3)
    ps = np.concatenate([sd[k].reshape(-1).float().numpy() for k in keys])[::TRAIN_STRIDE]
    # RMSprop's first steps move every weight by ~lr/sqrt(1-alpha) regardless of gradient size, so noise-level
    # gradients (the zero-gradient biases above) may step either way: compare within that step size
    assert np.abs(ps - g["param_sample"]).max() <= 2.1 * 2 * float(g["lr"]) / np.sqrt(1 - T.RMSPROP_ALPHA)
    assert np.median(np.abs(ps - g["param_sample"])) <= 1e-5
    bkeys = sorted(k for k in sd if k.endswith("running_mean") or k.endswith("running_var"))
    st = np.concatenate([sd[k].reshape(-1).numpy() for k in bkeys])[::7]
    np.testing.assert_allclose(st, g["stat_sample"], rtol=1e-4, atol=1e-5)
    assert int(sd["bn1.num_batches_tracked"]) == steps


def test_rmsprop_closed_form_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(50, requires_grad=True)
    opt = torch.optim.RMSprop([p], lr=1e-3, momentum=0, weight_decay=0)
    sd, state = {"p": p.detach().clone()}, {}
    for _ in range(3):
        g = torch.randn(50)
        p.grad = g.clone()
        opt.step()
        T.rmsprop_update(sd, {"p": g}, state, 1e-3)
    np.testing.assert_allclose(sd["p"].numpy(), p.detach().numpy(), rtol=1e-6, atol=1e-7)
    assert T.adjust_learning_rate(1.0, 3, [3, 5], 0.1) == 0.1 and T.adjust_learning_rate(1.0, 4, [3, 5], 0.1) == 1.0
