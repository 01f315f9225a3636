"""Host-side mirror of the reference's u_net_arch/pt_custom_ops/pt_utils.py: same public names,
constructor arguments, forward signatures and return values, running on the sm_100a library.

    grouping_operation(features, idx)                                   ref pt_utils.py:17-65
    masked_ordered_ball_query(radius, nsample, q_xyz, s_xyz, q_mask, s_mask)   ref :68-81
    masked_nearest_query(q_xyz, s_xyz, q_mask, s_mask)                   ref :84-96
    masked_grid_subsampling(xyz, mask, npoint, sampleDl)                 ref :99-112
    MaskedQueryAndGroup / MaskedNearestQueryAndGroup                     ref :115-180
    MaskedMaxPool / MaskedUpsample                                       ref :183-238

What differs from the reference is only HOW: neighbour lists are cached per forward (neighbors.py), and
MaskedMaxPool / MaskedUpsample('nearest') use the fused gather kernels instead of materialising the
(B, C, npoint, nsample) tensor.  The materialising groupers are kept for callers that need the grouped
tensors themselves (other aggregation operators, 'max'/'rbf' upsampling).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import _ext
from .. import neighbors
from ..fused import GatherMaxFunction, NearestGatherFunction


class GroupingOperation(Function):
    """(B, C, N) features, (B, npoint, nsample) int32 idx -> (B, C, npoint, nsample)."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.idx, ctx.n = idx, features.size(2)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        return _ext.group_points_grad(grad_out.contiguous(), ctx.idx, ctx.n), None


grouping_operation = GroupingOperation.apply


class MaskedOrderedBallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, query_xyz, support_xyz, query_mask, support_mask):
        nbr = neighbors.ball_neighbors(query_xyz, support_xyz, query_mask, support_mask, radius, nsample)
        # the reference hands out fresh tensors that callers clamp in place (pt_utils.py:126-127); the
        # cached list already holds only in-range indices, so handing out the cached tensors is equivalent
        ctx.mark_non_differentiable(nbr.idx, nbr.idx_mask)
        return nbr.idx, nbr.idx_mask

    @staticmethod
    def backward(ctx, *grads):
        return (None,) * 6


masked_ordered_ball_query = MaskedOrderedBallQuery.apply


class MaskedNearestQuery(Function):
    @staticmethod
    def forward(ctx, query_xyz, support_xyz, query_mask, support_mask):
        nbr = neighbors.nearest_neighbors(query_xyz, support_xyz, query_mask, support_mask)
        ctx.mark_non_differentiable(nbr.idx, nbr.idx_mask)
        return nbr.idx, nbr.idx_mask

    @staticmethod
    def backward(ctx, *grads):
        return (None,) * 4


masked_nearest_query = MaskedNearestQuery.apply


class MaskedGridSubsampling(Function):
    @staticmethod
    def forward(ctx, xyz, mask, npoint, sampleDl):
        sub_xyz, sub_mask = neighbors.grid_subsample(xyz, mask, npoint, sampleDl)
        ctx.mark_non_differentiable(sub_xyz, sub_mask)
        return sub_xyz, sub_mask

    @staticmethod
    def backward(ctx, *grads):
        return (None,) * 4


masked_grid_subsampling = MaskedGridSubsampling.apply


def _group(idx, query_xyz, support_xyz, features, radius, normalize_xyz, use_xyz):
    """Materialising gather shared by the two groupers (pt_utils.py:129-146 / :161-178)."""
    rel = grouping_operation(support_xyz.transpose(1, 2).contiguous(), idx)  # (B, 3, npoint, nsample)
    rel = rel - query_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        rel = rel / radius
    if features is None:
        assert use_xyz, "Cannot have not features and not use xyz as a feature!"
        return rel, rel
    grouped = grouping_operation(features.contiguous(), idx)  # channel-last views are re-laid out for the composed path
    return (torch.cat([rel, grouped], dim=1) if use_xyz else grouped), rel


class MaskedQueryAndGroup(nn.Module):
    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz

    def neighbors(self, query_xyz, support_xyz, query_mask, support_mask):
        """The cached NeighborList the fused operators consume (not part of the reference API)."""
        return neighbors.ball_neighbors(query_xyz, support_xyz, query_mask, support_mask, self.radius, self.nsample)

    def forward(self, query_xyz, support_xyz, query_mask, support_mask, features=None):
        idx, idx_mask = masked_ordered_ball_query(self.radius, self.nsample, query_xyz, support_xyz, query_mask,
                                                  support_mask)
        new_features, grouped_xyz = _group(idx, query_xyz, support_xyz, features, self.radius, self.normalize_xyz,
                                           self.use_xyz)
        if self.ret_grouped_xyz:
            return new_features, grouped_xyz, idx_mask
        return new_features, idx_mask


class MaskedNearestQueryAndGroup(nn.Module):
    def __init__(self, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz

    def forward(self, query_xyz, support_xyz, query_mask, support_mask, features=None):
        idx, idx_mask = masked_nearest_query(query_xyz, support_xyz, query_mask, support_mask)
        if self.normalize_xyz:
            # the reference reads an undefined self.radius here (pt_utils.py:164-165); fail the same way
            raise AttributeError("'MaskedNearestQueryAndGroup' object has no attribute 'radius'")
        new_features, grouped_xyz = _group(idx, query_xyz, support_xyz, features, None, False, self.use_xyz)
        if self.ret_grouped_xyz:
            return new_features, grouped_xyz, idx_mask
        return new_features, idx_mask


class MaskedMaxPool(nn.Module):
    def __init__(self, npoint, radius, nsample, sampleDl):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.sampleDl = npoint, radius, nsample, sampleDl
        self.grouper = MaskedQueryAndGroup(radius, nsample, use_xyz=False, ret_grouped_xyz=True)

    def forward(self, xyz, mask, features):
        sub_xyz, sub_mask = masked_grid_subsampling(xyz, mask, self.npoint, self.sampleDl)
        nbr = self.grouper.neighbors(sub_xyz, xyz, sub_mask, mask)
        sub_features = GatherMaxFunction.apply(features, nbr)  # max over all nsample slots, (B, C, npoint)
        return sub_xyz, sub_mask, sub_features


class MaskedUpsample(nn.Module):
    def __init__(self, radius, nsample, mode='nearest'):
        super().__init__()
        self.radius, self.nsample, self.mode = radius, nsample, mode
        if mode == 'nearest':
            self.grouper = MaskedNearestQueryAndGroup(use_xyz=False, ret_grouped_xyz=True)
        else:
            self.grouper = MaskedQueryAndGroup(radius, nsample, use_xyz=False, ret_grouped_xyz=True)

    def forward(self, up_xyz, xyz, up_mask, mask, features):
        if self.mode == 'nearest':
            nbr = neighbors.nearest_neighbors(up_xyz, xyz, up_mask, mask)
            return NearestGatherFunction.apply(features, nbr)
        grouped, grouped_xyz, _ = self.grouper(up_xyz, xyz, up_mask, mask, features)
        if self.mode == 'max':
            return F.max_pool2d(grouped, kernel_size=[1, grouped.shape[3]]).squeeze(-1)
        if self.mode == 'rbf':
            # weights exp(-|d|^2 / 2), normalised by nsample (pt_utils.py:230-234)
            rbf = torch.exp(-0.5 * grouped_xyz.pow(2).sum(1))
            return (grouped * rbf.unsqueeze(1)).sum(-1) / float(self.nsample)
        raise NotImplementedError(f"mode:{self.mode} not supported in MaskedUpsample")
