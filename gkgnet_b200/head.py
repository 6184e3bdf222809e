"""LabelQueryHead + its losses -- mirror of mmcls/models/heads/label_query_head.py.
On CUDA tensors the scores, both losses and their gradients run on the kernels of csrc/label_head.cu
(``ops.label_score`` / ``ops.multilabel_losses``); the PyTorch expressions below are the host-side definition the
CPU tests (gloo data-parallel plumbing, oracle comparison) exercise."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .registry import HEADS, register_into_mmcls


def asymmetric_loss(pred, target, gamma_pos=1.0, gamma_neg=4.0, clip=0.05, eps=1e-8):
    """Element-wise ASL on logits (losses/asymmetric_loss.py:9-72), summed over all entries."""
    p = pred.sigmoid()
    t = target.type_as(pred)
    if clip and clip > 0:
        pt = (1 - p + clip).clamp(max=1) * (1 - t) + p * t
    else:
        pt = (1 - p) * (1 - t) + p * t
    weight = (1 - pt).pow(gamma_pos * t + gamma_neg * (1 - t))
    return (-torch.log(pt.clamp(min=eps)) * weight).sum()


@HEADS.register_module()
class LabelQueryHead(nn.Module):
    """score[b, i] = fc1.weight[i] . L[b, i] + fc1.bias[i] + fc2(gap)[b, i]
    (label_query_head.py:49-57 computes the full (B, n, n) product and masks the diagonal)."""

    def __init__(self, num_classes, in_channels, softmax=False, double_loss=True,
                 init_cfg=dict(type="Normal", layer="Linear", std=0.01), loss=None, topk=(1,),
                 cal_acc=False, **kwargs):
        super().__init__()
        if num_classes <= 0:
            raise ValueError(f"num_classes={num_classes} must be a positive integer")
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.softmax = softmax
        self.double_loss = double_loss
        loss = dict(loss or dict(type="AsymmetricLoss", gamma_pos=0.0, gamma_neg=2.0, clip=0.05))
        loss.pop("type", None)
        self.loss_cfg = loss
        self.fc1 = nn.Linear(in_channels, num_classes)
        self.fc2 = nn.Linear(in_channels, num_classes)
        for lin in (self.fc1, self.fc2):
            nn.init.normal_(lin.weight, std=init_cfg.get("std", 0.01))
            nn.init.zeros_(lin.bias)

    def get_score(self, x):
        label_emb, gap = x[0], x[1]
        if label_emb.is_cuda:
            return ops.label_score(label_emb, gap, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)
        diag = (label_emb * self.fc1.weight.unsqueeze(0)).sum(-1) + self.fc1.bias
        return diag + self.fc2(gap)

    def simple_test(self, x, softmax=False, post_process=True):
        score = self.get_score(x)
        pred = F.softmax(score, dim=1) if self.softmax else torch.sigmoid(score)
        if post_process:
            return list(pred.detach().float().cpu().numpy())
        return pred

    def forward_train(self, x, gt_label, **kwargs):
        score = self.get_score(x).float()
        n = score.shape[0]
        if score.is_cuda:
            asl, bce = ops.multilabel_losses(score, gt_label, smooth=0.1, **self.loss_cfg)
            asl, bce = asl / n, bce / n
        else:
            asl = asymmetric_loss(score, gt_label, **self.loss_cfg) / n
            smooth = gt_label.type_as(score) * 0.8 + 0.1          # LabelSmoothLoss(0.1, 'multi_label')
            bce = F.binary_cross_entropy_with_logits(score, smooth, reduction="sum") / n
        if self.double_loss:
            return {"bce_loss": bce, "asy_loss": asl * 10.0}
        return {"loss": asl}


register_into_mmcls(LabelQueryHead)
