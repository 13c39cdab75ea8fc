"""Evaluator with the reference's structure (src/runner/evaluator.py:9-57); the loader is passed in
(datasets are out of scope)."""
import torch

from src.loss.mse import MSELoss
from src.utils.evaluation import AverageMeter, accuracy
from hgb200.ops import check_err_word


class Evaluator(object):
    def __init__(self, device, cfg, val_loader=None):
        self.cfg = cfg
        self.device = device
        self.criterion = MSELoss(use_target_weight=True)
        self.val_loader = val_loader
        if self.cfg['MODEL']['subset']:
            self.idxs = torch.LongTensor(cfg['MODEL']['subset'])

    def evaluate(self, model):
        model.eval()
        average_loss = AverageMeter()
        average_acc = AverageMeter()
        idxs = self.cfg['MODEL']['subset']
        with torch.no_grad():
            for i, (images, heatmaps, meta) in enumerate(self.val_loader):
                if idxs:
                    heatmaps = torch.index_select(heatmaps, 1, self.idxs)
                images = images.to(self.device)
                heatmaps = heatmaps.to(self.device, non_blocking=True)
                target_weight = meta['target_weight'].to(self.device, non_blocking=True)
                outputs = model(images)
                last_hms = outputs[-1]
                loss = self.criterion(outputs, heatmaps, target_weight)
                acc = accuracy(last_hms, heatmaps, idxs, thr=self.cfg['COMMON']['pck'])
                average_loss.update(loss.item(), images.size(0))
                check_err_word(heatmaps.device)
                average_acc.update(acc[0], images.size(0))
        return average_loss.avg, average_acc.avg
