"""In-situ checker for the training plan on the GPU (TEST INFRASTRUCTURE ONLY).

ShadowOps wraps hgb200.ops: every op the plan launches runs on the GPU as usual, and is ALSO replayed by the
CPU torch emulation (tests/fake_ops.py) on clones of the very same input buffers; all tensor arguments are
compared afterwards.  Because each launch is checked on the inputs the GPU actually produced, rounding noise
cannot accumulate across the network (train-mode BatchNorm amplifies it chaotically on a randomly
initialised model), so the tolerances stay at the level of one bf16 rounding while the check still covers
every launch, shape, aliasing pattern and buffer reuse of the real step."""
from __future__ import annotations

from collections import defaultdict

import torch

import fake_ops

PASSTHROUGH = {"halo_padded_buffer", "halo_interior", "halo_padded_elems", "gaussian_patch", "check_err_word", "err_word"}


class ShadowOps:
    def __init__(self, real_ops, bf16_tol=1.7e-2, f32_tol=5e-3):
        self.real = real_ops
        self.bf16_tol, self.f32_tol = bf16_tol, f32_tol
        self.stats = defaultdict(lambda: dict(calls=0, max_rel_l2=0.0, max_bad_frac=0.0))
        self.failures = []
        self._tables = {}

    # ---- helpers
    def _clone_tree(self, obj, memo):
        if torch.is_tensor(obj):
            key = (obj.data_ptr(), tuple(obj.shape), tuple(obj.stride()), obj.dtype)
            if key not in memo:
                memo[key] = (obj, obj.detach().cpu().clone())
            return memo[key][1]
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._clone_tree(o, memo) for o in obj)
        if isinstance(obj, dict):
            return {k: self._clone_tree(v, memo) for k, v in obj.items()}
        return obj

    def _compare(self, name, memo):
        st = self.stats[name]
        st["calls"] += 1
        for gpu_t, cpu_t in memo.values():
            a = gpu_t.detach().cpu().double()
            b = cpu_t.double()
            scale = float(b.abs().max())
            if scale == 0.0 and float(a.abs().max()) == 0.0:
                continue
            tol = self.bf16_tol if gpu_t.dtype == torch.bfloat16 else self.f32_tol
            diff = (a - b).abs()
            # a few elements may sit on a ReLU-mask / rounding boundary: judge the bulk (relative L2) and the
            # fraction of elements further than `tol` of the tensor's peak from the emulation
            rel_l2 = float(diff.norm() / (b.norm() + 1e-300))
            bad = float((diff > tol * max(scale, 1e-30)).double().mean())
            st["max_rel_l2"] = max(st["max_rel_l2"], rel_l2)
            st["max_bad_frac"] = max(st["max_bad_frac"], bad)
            # (short vectors: one element on a ReLU-mask boundary is already 0.4 % of a 256-channel sum vector)
            if rel_l2 > tol or bad > max(2e-3, 2.0 / a.numel()):
                self.failures.append(f"{name} #{st['calls']}: shape {tuple(gpu_t.shape)} {gpu_t.dtype} rel_l2 {rel_l2:.3e} "
                                     f"bad_frac {bad:.3e}")

    # ---- the pack table is a device byte blob for the real library and a list of dicts for the emulation
    def make_pack_table(self, entries, device):
        table = self.real.make_pack_table(entries, device)
        self._tables[table.data_ptr()] = entries
        return table

    def pack_weights(self, table, n_entries, reads=None, writes=None, which=3):
        entries = self._tables[table.data_ptr()]
        memo = {}
        cpu_entries = self._clone_tree(entries, memo)
        self.real.pack_weights(table, n_entries, which=which)
        torch.cuda.synchronize()
        fake_ops.pack_weights(cpu_entries, n_entries, which=which)
        self._compare("pack_weights", memo)

    def __getattr__(self, name):
        real_fn = getattr(self.real, name)
        if name in PASSTHROUGH or not callable(real_fn):
            return real_fn
        fake_fn = getattr(fake_ops, name)

        def wrapped(*args, **kw):
            memo = {}
            cargs = self._clone_tree(args, memo)
            ckw = self._clone_tree(kw, memo)
            out = real_fn(*args, **kw)
            torch.cuda.synchronize()
            fake_fn(*cargs, **ckw)
            self._compare(name, memo)
            return out

        return wrapped

    def report(self):
        lines = []
        for k, v in sorted(self.stats.items()):
            lines.append(f"{k:24s} calls {v['calls']:5d}  max rel L2 {v['max_rel_l2']:.3e}  max off-tolerance fraction {v['max_bad_frac']:.2e}")
        return "\n".join(lines)
