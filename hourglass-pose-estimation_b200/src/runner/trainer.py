"""Trainer with the reference's structure (src/runner/trainer.py:15-181) on the sm_100a training engine.

What changes against the reference: `torch.nn.DataParallel` (trainer.py:37: per-step parameter broadcast,
GIL-bound replica threads, loss/optimizer on GPU 0) becomes ONE PROCESS PER GPU under torchrun -- each rank
runs the fused step (hgb200.train.TrainEngine.train_step) on its batch shard, BatchNorm uses the shard's
statistics exactly as DataParallel replicas do, the flat gradient buffer is summed by ONE NCCL all-reduce
over NVLink (the local loss is pre-scaled by 1/world so the sum is the global-batch mean), and RMSprop is a
single fused launch over the flat parameter buffer.

Batches.  `cfg['TRAIN']['train_batch']` is the GLOBAL batch, as in the reference (DataParallel splits it over the
GPUs): every rank sees the same sequence of global batches and trains on its contiguous shard of each
(hgb200.shard.batch_shard), so the effective batch size and learning-rate regime do not change with the number of
GPUs.  The gradient scale is shard size / global batch size, so uneven shards of a ragged last batch still sum to
the global-batch mean, and a rank whose shard is empty still joins the all-reduce (TrainEngine.idle_step).
Loaders: pass any iterables yielding (images, heatmaps, {'target_weight': ...}); with none passed they are built
exactly as the reference builds them (trainer.py:47-58) from `src.datasets`, which re-exports the reference's
dataset classes when a reference checkout is available (src/datasets/__init__.py) -- the shuffling generator is
seeded identically on every rank so that all ranks draw the same global batches.
"""
import os

import torch
import torch.distributed as dist

from src.loss.mse import MSELoss
from src import models
from src.utils.evaluation import AverageMeter, accuracy
from hgb200 import ops
from hgb200.shard import batch_shard
from hgb200.train import train_engine, RMSPROP_ALPHA, RMSPROP_EPS
from hgb200.prefetch import DevicePrefetcher


def adjust_learning_rate(optimizer, epoch, lr, schedule, gamma):
    """Sets the learning rate to the initial LR decayed by schedule (trainer.py:15-21)."""
    if epoch in schedule:
        lr *= gamma
        for param_group in optimizer.param_groups:
            param_group['lr'] = lr
    return lr


class FusedRMSprop(object):
    """torch.optim.RMSprop(lr, momentum=0, weight_decay=0) (trainer.py:39-41) as one launch over the flat
    parameter / gradient / square-average buffers.  Exposes what the reference touches: param_groups (for
    adjust_learning_rate), zero_grad, step, state_dict / load_state_dict."""

    def __init__(self, engine, lr):
        self.engine = engine
        self.param_groups = [{'lr': lr, 'alpha': RMSPROP_ALPHA, 'eps': RMSPROP_EPS, 'momentum': 0, 'weight_decay': 0}]

    def zero_grad(self, set_to_none=False):
        pass                            # the step's backward pass starts by zeroing the flat gradient buffer

    def step(self):
        g = self.param_groups[0]
        self.engine.rmsprop(g['lr'], g['alpha'], g['eps'])

    def state_dict(self):
        """torch.optim.RMSprop's own state_dict layout (what the reference saves under 'optimizer', trainer.py:172,
        and feeds back through optimizer.load_state_dict on resume, trainer.py:71): per-parameter 'step' and
        'square_avg' indexed by position in model.parameters(), so checkpoints travel both ways."""
        st = self.engine.store
        names = list(st.slots)
        g = self.param_groups[0]
        group = {'lr': g['lr'], 'momentum': 0, 'alpha': g['alpha'], 'eps': g['eps'], 'centered': False, 'weight_decay': 0,
                 'capturable': False, 'foreach': None, 'maximize': False, 'differentiable': False,
                 'params': list(range(len(names)))}
        state = {}
        if self.engine.steps > 0:
            for i, k in enumerate(names):
                state[i] = {'step': torch.tensor(float(self.engine.steps)),
                            'square_avg': st.view(st.V, k).detach().clone().contiguous()}
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        st = self.engine.store
        names = list(st.slots)
        if 'square_avg' in sd:                       # this build's earlier layout: {name: tensor}
            for k, v in sd['square_avg'].items():
                st.view(st.V, k).copy_(v)
            self.param_groups[0].update({k: v for k, v in sd['param_groups'][0].items() if k in ('lr', 'alpha', 'eps')})
            return
        group = sd['param_groups'][0]
        if group.get('momentum', 0) != 0 or group.get('centered', False) or group.get('weight_decay', 0) != 0:
            raise ValueError("FusedRMSprop implements the reference's configuration only: momentum=0, centered=False, "
                             "weight_decay=0 (trainer.py:39-41)")
        self.param_groups[0].update(lr=group['lr'], alpha=group.get('alpha', RMSPROP_ALPHA), eps=group.get('eps', RMSPROP_EPS))
        order = group['params']
        if len(order) != len(names):
            raise ValueError(f"optimizer state has {len(order)} parameters, the model has {len(names)}")
        steps = 0
        for pos, idx in enumerate(order):
            ent = sd['state'].get(idx)
            if ent is None:
                st.view(st.V, names[pos]).zero_()
                continue
            st.view(st.V, names[pos]).copy_(ent['square_avg'])
            steps = max(steps, int(float(ent.get('step', 0))))
        self.engine.steps = steps


class Trainer(object):
    def __init__(self, cfg, num_classes, train_loader=None, val_loader=None):
        self.cfg = cfg
        self.rank = int(os.environ.get('RANK', 0))
        self.world = int(os.environ.get('WORLD_SIZE', 1))
        local_rank = int(os.environ.get('LOCAL_RANK', 0))
        if not torch.cuda.is_available():
            raise RuntimeError("Trainer (B200 build) needs a CUDA device: there is no CPU fallback")
        self.device = torch.device('cuda', local_rank)
        torch.cuda.set_device(self.device)
        self._own_pg = False
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group(backend='nccl', device_id=self.device)
            self._own_pg = True
        if self.rank == 0:
            print(f"==> creating model '{cfg['MODEL']['arch']}', stacks={cfg['MODEL']['num_stacks']}")
        torch.manual_seed(cfg.get('COMMON', {}).get('seed', 0))      # identical initial weights on every rank
        model = models.__dict__[cfg['MODEL']['arch']](num_stacks=cfg['MODEL']['num_stacks'],
                                                      num_blocks=1,
                                                      num_classes=num_classes,
                                                      mobile=cfg['MODEL']['mobile'],
                                                      skip_mode=cfg['MODEL']['skip_mode'],
                                                      out_res=cfg['DATASET']['out_res'])
        self.model = model.to(self.device)
        self.engine = train_engine(self.model)
        self.optimizer = FusedRMSprop(self.engine, cfg['TRAIN']['learning_rate'])
        self.criterion = MSELoss(use_target_weight=True)
        self.start_epoch = 0
        self.best_acc = 0
        self.train_loader = train_loader
        self.val_loader = val_loader
        if train_loader is None or val_loader is None:
            self._build_loaders(train_loader is None, val_loader is None)
        self.idxs = cfg['MODEL']['subset']
        if os.path.isfile(cfg['COMMON'].get('resume', '') or ''):
            self._resume()

    def _build_loaders(self, want_train, want_val):
        """trainer.py:47-58: the reference's datasets and DataLoaders (global batches; same shuffle on every rank)."""
        from src import datasets
        import torch.utils.data
        cfg = self.cfg
        name = cfg['DATASET']['name']
        if name not in datasets.__dict__:
            raise KeyError(f"dataset '{name}' is not available: src.datasets re-exports the reference's dataset classes "
                           f"from a reference checkout (HG_REFERENCE_SRC=<checkout>/src); pass train_loader / val_loader "
                           f"otherwise")
        seed = cfg.get('COMMON', {}).get('seed', 0)
        if want_train:
            train_dataset = datasets.__dict__[name](is_train=True, **cfg['DATASET'])
            self.train_loader = torch.utils.data.DataLoader(
                train_dataset, batch_size=cfg['TRAIN']['train_batch'], shuffle=True,
                num_workers=cfg['TRAIN']['num_workers'], pin_memory=True,
                generator=torch.Generator().manual_seed(seed))
        if want_val:
            val_dataset = datasets.__dict__[name](is_train=False, **cfg['DATASET'])
            self.val_loader = torch.utils.data.DataLoader(
                val_dataset, batch_size=cfg['TRAIN']['val_batch'], shuffle=True,
                num_workers=cfg['TRAIN']['num_workers'], pin_memory=True,
                generator=torch.Generator().manual_seed(seed + 1))

    # ------------------------------------------------------------------ checkpoints (reference format)
    def _resume(self):
        # reference checkpoints are plain pickles (numpy scalars in 'best_acc'); same trust model as the reference's torch.load
        checkpoint = torch.load(self.cfg['COMMON']['resume'], map_location=self.device, weights_only=False)
        self.start_epoch = checkpoint['epoch']
        self.best_acc = float(checkpoint['best_acc'])
        sd = {(k[7:] if k.startswith('module.') else k): v for k, v in checkpoint['state_dict'].items()}
        self.model.load_state_dict(sd)
        if checkpoint.get('optimizer'):
            self.optimizer.load_state_dict(checkpoint['optimizer'])

    def state(self, epoch):
        # keys carry the 'module.' prefix the reference's DataParallel wrapper adds (estimator.py:28-31 strips it)
        return {'epoch': epoch + 1,
                'state_dict': {'module.' + k: v.detach().clone().contiguous() for k, v in self.model.state_dict().items()},
                'optimizer': self.optimizer.state_dict(),
                'best_acc': float(self.best_acc)}

    # ------------------------------------------------------------------ one step / one epoch
    def _all_reduce(self, flat_grads):
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)

    def train_step(self, images, heatmaps, target_weight, global_batch=None):
        """trainer.py:82-99 for THIS rank's shard of a global batch of `global_batch` images (default: the shard is one
        of `world` equal ones).  Returns (loss device tensor [1], last heat maps)."""
        lr = self.optimizer.param_groups[0]['lr']
        n, _, h, w = images.shape
        scale = n / float(global_batch) if global_batch else 1.0 / self.world
        loss = self.engine.train_step(images.to(self.device, non_blocking=True),
                                      heatmaps.to(self.device, non_blocking=True),
                                      target_weight.to(self.device, non_blocking=True), lr,
                                      world_size=self.world, grad_scale=scale,
                                      all_reduce=self._all_reduce if self.world > 1 else None)
        return loss, self.engine.plans[(n, h, w)].outputs[-1]

    def _train_epoch(self):
        self.model.train()
        average_loss = AverageMeter()
        average_acc = AverageMeter()
        sizes = []                       # global batch size of every step, in order (read when the shard arrives)

        def host_batches():
            for images, heatmaps, meta in self.train_loader:
                if self.idxs:
                    heatmaps = torch.index_select(heatmaps, 1, torch.LongTensor(self.idxs))
                # this rank's contiguous shard of the global batch (the reference's DataParallel scatter, trainer.py:37)
                lo, hi = batch_shard(images.size(0), self.world, self.rank)
                sizes.append(images.size(0))
                yield images[lo:hi], heatmaps[lo:hi], meta['target_weight'][lo:hi]

        # the batch of step i+1 crosses PCIe on a side stream while step i computes (hgb200/prefetch.py)
        for step, (images, heatmaps, target_weight) in enumerate(DevicePrefetcher(host_batches(), self.device)):
            if images.size(0) == 0:      # fewer images than ranks in a ragged last batch: still join the collective
                self.engine.idle_step(self.optimizer.param_groups[0]['lr'],
                                      self._all_reduce if self.world > 1 else None)
                continue
            loss, last_hms = self.train_step(images, heatmaps, target_weight, global_batch=sizes[step])
            acc = accuracy(last_hms, heatmaps, self.idxs, thr=self.cfg['COMMON']['pck'])
            average_loss.update(loss.item(), images.size(0))
            ops.check_err_word(self.device)          # the host has just synchronised: a kernel-side protocol timeout surfaces here
            average_acc.update(acc[0], images.size(0))
        return average_loss.avg, average_acc.avg

    def _evaluate(self):
        self.model.eval()
        average_loss = AverageMeter()
        average_acc = AverageMeter()
        with torch.no_grad():
            for i, (images, heatmaps, meta) in enumerate(self.val_loader):
                if self.idxs:
                    heatmaps = torch.index_select(heatmaps, 1, torch.LongTensor(self.idxs))
                images = images.to(self.device)
                heatmaps = heatmaps.to(self.device, non_blocking=True)
                target_weight = meta['target_weight'].to(self.device, non_blocking=True)
                outputs = self.model(images)
                last_hms = outputs[-1]
                loss = self.criterion(outputs, heatmaps, target_weight)
                acc = accuracy(last_hms, heatmaps, self.idxs, thr=self.cfg['COMMON']['pck'])
                average_loss.update(loss.item(), images.size(0))
                ops.check_err_word(self.device)
                average_acc.update(acc[0], images.size(0))
        is_best = False
        if average_acc.avg > self.best_acc:
            is_best = True
            self.best_acc = average_acc.avg
        return average_loss.avg, average_acc.avg, is_best

    def close(self):
        """Release the captured CUDA graphs and, if this Trainer created it, the process group -- in that order: a graph
        that holds the all-reduce nodes (HG_OVERLAP_AR=1) must be destroyed before the NCCL communicator is."""
        self.engine.release_graphs()
        if self._own_pg and dist.is_initialized():
            dist.destroy_process_group()
            self._own_pg = False

    def train(self):
        only_checkpoint_path = os.path.join(self.cfg['COMMON']['checkpoint_dir'], 'ckpts')
        if self.rank == 0 and not os.path.isdir(only_checkpoint_path):
            os.makedirs(only_checkpoint_path)
        lr = self.cfg['TRAIN']['learning_rate']
        for epoch in range(self.start_epoch, self.cfg['TRAIN']['epochs'] + 1):
            lr = adjust_learning_rate(self.optimizer, epoch, lr, self.cfg['TRAIN']['schedule'],
                                      self.cfg['TRAIN']['gamma'])
            if self.rank == 0:
                print('\nEpoch: %d | LR: %.8f' % (epoch + 1, lr))
            loss, acc = self._train_epoch()
            val_loss, val_acc, is_best = self._evaluate()
            if self.rank == 0 and ((epoch + 1) % self.cfg['COMMON']['snapshot'] == 0 or is_best):
                state = self.state(epoch)
                if (epoch + 1) % self.cfg['COMMON']['snapshot'] == 0:
                    torch.save(state, os.path.join(only_checkpoint_path, f'checkpoint_{epoch+1}.pth.tar'))
                if is_best:
                    torch.save(state, os.path.join(only_checkpoint_path, 'best.pth.tar'))
        self.engine.release_graphs()      # re-captured on demand; nothing that holds NCCL nodes outlives the process group
