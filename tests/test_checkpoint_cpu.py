"""Checkpoint compatibility (reference: src/runner/trainer.py:66-73,166-181): the optimizer state
this build saves is torch.optim.RMSprop's own state_dict layout, so a reference checkpoint resumes here and a
checkpoint written here resumes under the reference's `self.optimizer.load_state_dict(checkpoint['optimizer'])`."""
import io
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fake_ops  # noqa: E402
from oracle.hourglass_oracle import make_state_dict  # noqa: E402
from oracle.make_golden_inputs import train_inputs  # noqa: E402


@pytest.fixture()
def cpu_train(monkeypatch):
    import hgb200.train as tr
    monkeypatch.setattr(tr, "ops", fake_ops)
    return tr


def _model(S=1, J=16):
    from src.models import hg
    sd = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0)
    model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum")
    model.load_state_dict(sd)
    return model.train()


def test_optimizer_state_round_trips_with_torch_rmsprop(cpu_train):
    from src.runner.trainer import FusedRMSprop
    lr = 2.5e-4
    model = _model()
    eng = cpu_train.TrainEngine(model, "cpu")
    opt = FusedRMSprop(eng, lr)
    assert opt.state_dict()["state"] == {}                     # like a fresh torch optimizer
    x, tg, tw = train_inputs(1, 2, 16, 64, 64, 1)[0]
    eng.train_step(x, tg, tw, lr, use_graph=False)             # includes the fused RMSprop update
    sd = opt.state_dict()
    # -> the reference's optimizer accepts it (same parameter order as model.parameters())
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in model.parameters()]
    ref_opt = torch.optim.RMSprop(ref_params, lr=lr, momentum=0, weight_decay=0)
    buf = io.BytesIO()
    torch.save({"epoch": 1, "state_dict": {"module." + k: v for k, v in model.state_dict().items()}, "optimizer": sd,
                "best_acc": 0.5}, buf)
    buf.seek(0)
    ck = torch.load(buf)
    ref_opt.load_state_dict(ck["optimizer"])
    for p_ref, (name, p) in zip(ref_params, model.named_parameters()):
        sq = ref_opt.state[p_ref]["square_avg"]
        assert torch.equal(sq, eng.store.view(eng.store.V, name))
        assert float(ref_opt.state[p_ref]["step"]) == 1.0
    assert ref_opt.param_groups[0]["lr"] == lr and ref_opt.param_groups[0]["alpha"] == 0.99
    # one more step on identical gradients gives identical parameters under both optimizers
    for p_ref, p in zip(ref_params, model.parameters()):
        p_ref.grad = p.grad.detach().clone().contiguous()
    ref_opt.step()
    opt.step()
    for p_ref, p in zip(ref_params, model.parameters()):
        assert float((p_ref.detach() - p.detach()).abs().max()) <= 1e-6
    # <- and a state written by torch.optim.RMSprop loads here
    model2 = _model()
    eng2 = cpu_train.TrainEngine(model2, "cpu")
    opt2 = FusedRMSprop(eng2, 1.0)
    opt2.load_state_dict(ref_opt.state_dict())
    assert opt2.param_groups[0]["lr"] == lr and eng2.steps == 2
    for p_ref, (name, _) in zip(ref_params, model2.named_parameters()):
        assert torch.equal(ref_opt.state[p_ref]["square_avg"], eng2.store.view(eng2.store.V, name))
    # only the reference's RMSprop configuration is implemented: anything else is refused, not approximated
    bad = ref_opt.state_dict()
    bad["param_groups"][0]["momentum"] = 0.9
    with pytest.raises(ValueError):
        opt2.load_state_dict(bad)


def test_module_prefixed_state_dict_loads(cpu_train):
    """DataParallel's 'module.' prefix (reference checkpoints; estimator.py:28-31 strips seven characters)."""
    model = _model()
    sd = {"module." + k: v.clone() for k, v in model.state_dict().items()}
    stripped = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    model2 = _model()
    with torch.no_grad():
        for p in model2.parameters():
            p.add_(1.0)
    model2.load_state_dict(stripped, strict=True)
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
