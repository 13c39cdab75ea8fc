"""The runner mirrors end to end on one GPU (reference: src/runner/trainer.py, evaluator.py, estimator.py): Trainer.train()
over synthetic loaders with the reference's cfg layout, checkpoint files in the reference's format (module.-prefixed
state_dict, torch.optim.RMSprop-shaped optimizer state), resume, Evaluator, and the shipped YAML's mobile=True."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_trained_accuracy import synthetic_batch, J, H, W  # noqa: E402

pytestmark = pytest.mark.gpu


def _cfg(tmp, mobile, epochs=1):
    return {"MODEL": {"arch": "hg", "num_stacks": 2, "mobile": mobile, "skip_mode": "sum", "subset": None},
            "DATASET": {"out_res": H // 4, "inp_res": H},
            "TRAIN": {"learning_rate": 2.5e-4, "epochs": epochs, "schedule": [1], "gamma": 0.1},
            "COMMON": {"checkpoint_dir": str(tmp), "snapshot": 1, "pck": 0.5, "resume": "", "seed": 0}}


def _loader(n_batches, batch, seed):
    """What the reference's DataLoader yields: (images, heat maps, {'target_weight': [B,J,1]}) on the host."""
    from oracle import loss_oracle as L
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        img, joints, vis = synthetic_batch(batch, gen)
        tg, tw = [], []
        for b in range(batch):
            t, w = L.generate_target(joints[b].numpy(), vis[b].numpy(), (W, H), (W // 4, H // 4), 1)
            tg.append(t)
            tw.append(w)
        out.append((img, torch.from_numpy(np.stack(tg)), {"target_weight": torch.from_numpy(np.stack(tw))}))
    return out


@pytest.mark.parametrize("mobile", [False, True])
def test_trainer_trains_checkpoints_resumes_and_evaluates(tmp_path, mobile):
    from src.runner.trainer import Trainer
    from src.runner.evaluator import Evaluator
    train, val = _loader(12, 8, 1), _loader(2, 8, 2)
    t = Trainer(_cfg(tmp_path, mobile), J, train_loader=train, val_loader=val)
    first_loss = t._train_epoch()[0]
    t.train()                                   # epochs 0..1: LR decays at epoch 1 (schedule), snapshot every epoch
    assert abs(t.optimizer.param_groups[0]["lr"] - 2.5e-5) < 1e-12
    ck = os.path.join(str(tmp_path), "ckpts", "checkpoint_2.pth.tar")
    assert os.path.isfile(ck) and os.path.isfile(os.path.join(str(tmp_path), "ckpts", "best.pth.tar"))
    state = torch.load(ck, map_location="cpu")
    assert set(state) == {"epoch", "state_dict", "optimizer", "best_acc"} and state["epoch"] == 2
    assert all(k.startswith("module.") for k in state["state_dict"])
    assert set(state["optimizer"]) == {"state", "param_groups"}
    # the reference's own optimizer accepts the saved state
    ref_params = [torch.nn.Parameter(v.clone()) for k, v in state["state_dict"].items()
                  if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]
    torch.optim.RMSprop(ref_params, lr=1.0).load_state_dict(state["optimizer"])
    val_loss, val_acc, _ = t._evaluate()
    assert val_loss < first_loss and np.isfinite(val_acc)
    # resume: a fresh Trainer picks up epoch, best accuracy, weights and optimizer state
    cfg2 = _cfg(tmp_path, mobile)
    cfg2["COMMON"]["resume"] = ck
    t2 = Trainer(cfg2, J, train_loader=train, val_loader=val)
    assert t2.start_epoch == 2 and t2.best_acc == state["best_acc"]
    for (k, a), (_, b) in zip(t.model.state_dict().items(), t2.model.state_dict().items()):
        assert torch.equal(a, b), k
    assert torch.equal(t.engine.store.V, t2.engine.store.V)
    ev = Evaluator(torch.device("cuda"), cfg2, val_loader=val)
    loss_e, acc_e = ev.evaluate(t2.model)
    assert abs(loss_e - val_loss) <= 1e-3 * val_loss and abs(acc_e - val_acc) < 1e-6
