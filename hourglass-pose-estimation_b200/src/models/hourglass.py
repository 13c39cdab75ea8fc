"""HourglassNet with the reference's constructor / forward signatures and state_dict keys
(reference: src/models/hourglass.py:7-97), executed by hand-written sm_100a kernels.

forward(x: fp32 NCHW [B,3,H,W]) -> python list of num_stacks fp32 tensors [B,num_classes,H/4,W/4],
exactly what the reference's runners consume (trainer.py:89-91, estimator.py:88).
"""
from operator import attrgetter

import torch
import torch.nn as nn

from src.models.modules import Hourglass, HGBottleneck

__all__ = ['HourglassNet', 'hg']

_VERSION = attrgetter('_version')


class HourglassNet(nn.Module):
    """Hourglass model from Newell et al ECCV 2016"""

    def __init__(self, block, num_stacks=2, num_blocks=4,
                 num_classes=16, mobile=False, skip_mode='sum'):
        super(HourglassNet, self).__init__()

        self.mobile = mobile
        self.inplanes = 64
        self.num_feats = 128
        self.num_stacks = num_stacks
        self.conv1 = nn.Conv2d(3, self.inplanes, kernel_size=7, stride=2, padding=3,
                               bias=True)
        self.bn1 = nn.BatchNorm2d(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_residual(block, self.inplanes, 1)
        self.layer2 = self._make_residual(block, self.inplanes, 1)
        self.layer3 = self._make_residual(block, self.num_feats, 1)
        self.maxpool = nn.MaxPool2d(2, stride=2)

        # build hourglass modules
        ch = self.num_feats * block.expansion
        hg, res, fc, score, fc_, score_ = [], [], [], [], [], []
        for i in range(num_stacks):
            hg.append(Hourglass(block, num_blocks, self.num_feats, 4,
                                mobile=self.mobile, skip_mode=skip_mode))
            res.append(self._make_residual(block, self.num_feats, num_blocks))
            fc.append(self._make_fc(ch, ch))
            score.append(nn.Conv2d(ch, num_classes, kernel_size=1, bias=True))
            if i < num_stacks - 1:
                fc_.append(nn.Conv2d(ch, ch, kernel_size=1, bias=True))
                score_.append(nn.Conv2d(num_classes, ch, kernel_size=1, bias=True))
        self.hg = nn.ModuleList(hg)
        self.res = nn.ModuleList(res)
        self.fc = nn.ModuleList(fc)
        self.score = nn.ModuleList(score)
        self.fc_ = nn.ModuleList(fc_)
        self.score_ = nn.ModuleList(score_)

        # engine state (not part of the state_dict)
        self._engine = None
        self._engine_key = None
        self.use_cuda_graph = True

    def _make_residual(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion,
                          kernel_size=1, stride=stride, bias=True),
            )

        layers = [block(self.inplanes, planes, stride, downsample, self.mobile)]
        self.inplanes = planes * block.expansion
        for i in range(1, blocks):
            layers.append(block(self.inplanes, planes))

        return nn.Sequential(*layers)

    def _make_fc(self, inplanes, outplanes):
        bn = nn.BatchNorm2d(inplanes)
        conv = nn.Conv2d(inplanes, outplanes, kernel_size=1, bias=True)
        return nn.Sequential(
            conv,
            bn,
            self.relu,
        )

    # ------------------------------------------------------------------ sm_100a execution
    def _weights_key(self, device):
        # in-place updates (optimizer.step, load_state_dict's copy_) bump tensor._version
        # (the fused training step updates parameters through raw pointers and bumps _weights_epoch instead).
        # Read straight from the leaf modules' own dictionaries: state_dict() / parameters() rebuild name prefixes and walk
        # the module tree through generators on every call -- 1.4 ms per forward for the 8-stack network, more than the
        # batch-1 forward's GPU time (scripts/estimate.py's case).  The module LIST is cached (this network never grows
        # submodules after construction); the tensors are looked up each time, so a replaced Parameter is still seen.
        leaves = self.__dict__.get("_key_leaves")
        if leaves is None:
            leaves = [(m._parameters, m._buffers) for m in self.modules() if m._parameters or m._buffers]
            self.__dict__["_key_leaves"] = leaves
        ps = [t for pd, _ in leaves for t in pd.values() if t is not None]
        bs = [t for _, bd in leaves for t in bd.values() if t is not None]
        version = sum(map(_VERSION, ps)) + sum(map(_VERSION, bs))
        ident = sum(map(id, ps))
        return (str(device), version, ident, getattr(self, "_weights_epoch", 0))

    def engine(self, device=None):
        """The folded-weight inference engine for the current parameters (rebuilt when they change)."""
        from hgb200.engine import HourglassEngine
        device = device or next(self.parameters()).device
        key = self._weights_key(device)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None and self._engine.device == torch.device(device):
                self._engine.update_weights(self.state_dict())         # same module, new values: keep plans and graphs
            else:
                self._engine = HourglassEngine(self.state_dict(), device)
            self._engine_key = key
        return self._engine

    def forward(self, x):
        if self.training:
            from hgb200.train import training_forward
            return training_forward(self, x)
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError("HourglassNet (B200 build) runs on CUDA only: there is no CPU fallback; "
                               "move the model with .to('cuda')")
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        return self.engine(dev).forward(x, use_graph=self.use_cuda_graph)


def hg(**kwargs):
    model = HourglassNet(HGBottleneck, num_stacks=kwargs['num_stacks'],
                         num_blocks=kwargs['num_blocks'], num_classes=kwargs['num_classes'],
                         mobile=kwargs['mobile'], skip_mode=kwargs['skip_mode'])
    return model
