"""MaskedL1Loss (ref u_net_arch/models/losses/masked_l1_loss.py:6-14): per-point mean absolute error over
the 3 offset components, padded points masked out, averaged over the valid points."""
import torch.nn as nn


class MaskedL1Loss(nn.Module):
    def forward(self, pred, target, mask):
        per_point = (pred - target).abs().mean(2) * mask
        return per_point.sum() / mask.sum()
