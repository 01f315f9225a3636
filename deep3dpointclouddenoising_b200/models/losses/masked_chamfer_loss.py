"""Chamfer training losses on the masked nearest-query kernel.

Mirrors of the reference modules (same constructor arguments, same forward(pred, target, mask, points)):
  ref: u_net_arch/models/losses/masked_chamfer_loss.py:10-30            MaskedChamferLoss
  ref: u_net_arch/models/losses/masked_chamfer_l1_loss.py:10-53         MaskedChamferL1Loss
  ref: u_net_arch/models/losses/masked_adaptive_l1_chamfer_loss.py:10-57 MaskedAdaptiveL1ChamferLoss
  ref: u_net_arch/models/losses/chamfer_distance_aux.py:153-246          chamfer_distance (K = 1 nearest neighbours in both
       directions by pytorch3d.knn_points, 'L2' = squared distance, 'L1' = sum of absolute differences to the nearest
       point, point reduction mean, then summed over the two directions)
The reference loops over the batch and calls pytorch3d on the valid prefix of every patch; here ONE launch of
d3d_nearest_query per direction handles the whole batch (the valid points are a prefix, which is the kernel's mask
convention), and the distances are re-evaluated with torch so that autograd differentiates them w.r.t. both clouds with the
neighbour indices held fixed — what knn_points' backward does."""
import torch
import torch.nn as nn

from ... import ops


def _nearest_rows(a, b, mask):
    """b's row nearest to every row of a, per batch element, restricted to the valid prefix: (B, N, 3).
    d3d_nearest_query follows the reference kernel (masked_nearest_query_gpu.cu:36-43): it stops at the first 0 of the
    mask and only accepts squared distances below 100; the reference LOSS has neither restriction (boolean mask
    selection + pytorch3d knn), so both are checked here instead of silently returning a different loss."""
    with torch.no_grad():
        idx, _ = ops.nearest_query(a.detach().contiguous(), b.detach().contiguous(), mask, mask)
        idx = idx.view(a.shape[0], a.shape[1])
        valid = mask != 0
        if bool(((idx < 0) & valid).any()):
            raise ValueError("chamfer loss: a valid point has no neighbour within squared distance 100 "
                             "(masked_nearest_query's cap); rescale the clouds to the unit-size convention")
        idx = idx.clamp(min=0).long()
    return torch.gather(b, 1, idx.unsqueeze(-1).expand(-1, -1, 3))


def _check_prefix_mask(mask):
    """The batched nearest query treats the mask as a valid PREFIX (the dataset guarantees it, offset_dataset.py:671-672)."""
    m = mask != 0
    if bool((m[:, 1:] & ~m[:, :-1]).any()):
        raise ValueError("chamfer loss: mask must be a valid prefix (ones then zeros) for the batched nearest query")


def masked_chamfer(clean_points, pred_points, mask, norm_type="L2"):
    """Mean over the batch of chamfer_distance(clean[valid], pred[valid], point_reduction='mean')."""
    if norm_type not in ("L2", "L1"):
        raise ValueError(f"Norm type {norm_type} not implemented")
    _check_prefix_mask(mask)
    mask_i = mask.int().contiguous()
    w = mask.to(clean_points.dtype)
    total = 0
    for a, b in ((clean_points, pred_points), (pred_points, clean_points)):
        diff = a - _nearest_rows(a, b, mask_i)
        d = (diff * diff).sum(2) if norm_type == "L2" else diff.abs().sum(2)
        total = total + ((d * w).sum(1) / w.sum(1)).sum()
    return total / mask.shape[0]


def _masked_l1(pred, target, mask):
    per_point = (pred - target).abs().mean(2) * mask
    return per_point.sum() / mask.sum()


class MaskedChamferLoss(nn.Module):
    def __init__(self, norm_type="L2"):
        super().__init__()
        self.norm_type = norm_type

    def forward(self, pred, target, mask, points):
        return masked_chamfer(points + target, points + pred, mask, self.norm_type)


class MaskedChamferL1Loss(nn.Module):
    def __init__(self, norm_type="L2"):
        super().__init__()
        self.norm_type = norm_type

    def forward(self, pred, target, mask, points):
        return 0.5 * (_masked_l1(pred, target, mask) + masked_chamfer(points + target, points + pred, mask, self.norm_type))


class MaskedAdaptiveL1ChamferLoss(nn.Module):
    def __init__(self, converging_to):
        super().__init__()
        self.converging_to = converging_to

    def forward(self, pred, target, mask, points):
        l1 = _masked_l1(pred, target, mask)
        cd = masked_chamfer(points + target, points + pred, mask, "L1")  # comparable to the L1 term (ref :33)
        if self.converging_to == 'chamfer':
            return l1 + torch.exp(-l1) * cd
        if self.converging_to == 'L1':
            return cd + torch.exp(-cd) * l1
        raise ValueError(f"Limit of loss {self.converging_to} not implemented")
