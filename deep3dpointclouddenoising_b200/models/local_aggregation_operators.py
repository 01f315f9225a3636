"""Local aggregation operators of the point U-Net on the fused sm_100a kernels.

Mirrors the public surface of u_net_arch/models/local_aggregation_operators.py for the operators on the
hot path: PosPool (:94-190), PseudoGrid (:403-510) and the LocalAggregation dispatcher (:513-551) — same
constructor signature (in_channels, out_channels, radius, nsample, config), same forward signature, same
parameter / buffer names (out_transform.*, out_conv.*, kernel_weights, K_points), so reference
checkpoints load.  AdaptiveWeight / PointWiseMLP / Attention are outside the path.

Where the reference gathers a (B, C, npoint, nsample) tensor and runs eager PyTorch on it, these modules
call one fused kernel (fused.py).  Variants the fused kernels do not cover (PosPool 'sin_cos' embedding,
'max' reduction) run the reference's formula on the materialised gather from our group_points kernel.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..fused import PosPoolFunction, PseudoGridFunction
from .blocks import conv_bn
from ..pt_custom_ops.pt_utils import MaskedQueryAndGroup
from ..utils.config import runtime
from .utlis import create_kernel_points, weight_variable


def _output_block(in_channels, out_channels, momentum, with_conv):
    return conv_bn(in_channels, out_channels, momentum, relu=True, conv=with_conv)


class PosPool(nn.Module):
    def __init__(self, in_channels, out_channels, radius, nsample, config):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.radius, self.nsample = radius, nsample
        self.position_embedding = config.pospool.position_embedding
        self.reduction = config.pospool.reduction
        self.output_conv = config.pospool.output_conv or (in_channels != out_channels)
        self.grouper = MaskedQueryAndGroup(radius, nsample, use_xyz=False, ret_grouped_xyz=True, normalize_xyz=True)
        block = _output_block(in_channels, out_channels, config.bn_momentum, self.output_conv)
        if self.output_conv:
            self.out_conv = block
        else:
            self.out_transform = block

    def _fused_ok(self, channels):
        return (self.position_embedding == 'xyz' and self.reduction in ('sum', 'avg', 'mean')
                and channels % 3 == 0 and channels % 4 == 0)

    def _composed(self, query_xyz, support_xyz, query_mask, support_mask, features):
        """Reference formula on the materialised gather (:140-183) for the variants without a fused kernel."""
        B, C, npoint = features.shape[0], features.shape[1], query_xyz.shape[1]
        grouped, rel, nmask = self.grouper(query_xyz, support_xyz, query_mask, support_mask, features)
        if self.position_embedding == 'xyz':
            agg = (rel.unsqueeze(1) * grouped.view(B, C // 3, 3, npoint, self.nsample)).view(B, C, npoint, self.nsample)
        elif self.position_embedding == 'sin_cos':
            feat_dim = C // 6
            freq = torch.pow(1000.0, torch.arange(feat_dim, dtype=torch.float32, device=rel.device) / feat_dim)
            ang = (100 * rel).unsqueeze(-1) / freq  # (B, 3, npoint, nsample, feat_dim)
            emb = torch.cat([ang.sin(), ang.cos()], -1).permute(0, 1, 4, 2, 3).reshape(B, C, npoint, self.nsample)
            agg = grouped * emb
        else:
            raise NotImplementedError(f'Position Embedding {self.position_embedding} not implemented in PosPool')
        if self.reduction == 'max':
            return F.max_pool2d(agg, kernel_size=[1, self.nsample]).squeeze(-1)
        if self.reduction in ('avg', 'mean', 'sum'):
            fmask = (nmask + (1 - query_mask[:, :, None]))[:, None]
            out = (agg * fmask).sum(-1)
            return out / fmask.sum(-1) if self.reduction != 'sum' else out
        raise NotImplementedError(f'Reduction {self.reduction} not implemented in PosPool ')

    def forward(self, query_xyz, support_xyz, query_mask, support_mask, support_features):
        if self._fused_ok(support_features.shape[1]):
            nbr = self.grouper.neighbors(query_xyz, support_xyz, query_mask, support_mask)
            out = PosPoolFunction.apply(support_features, query_xyz, support_xyz, query_mask, nbr, self.radius,
                                        'sum' if self.reduction == 'sum' else 'avg')
        else:
            out = self._composed(query_xyz, support_xyz, query_mask, support_mask, support_features)
        return self.out_conv(out) if self.output_conv else self.out_transform(out)


class PseudoGrid(nn.Module):
    def __init__(self, in_channels, out_channels, radius, nsample, config):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.radius, self.nsample = radius, nsample
        pg = config.pseudo_grid
        self.KP_influence = pg.KP_influence
        self.num_kernel_points = pg.num_kernel_points
        self.convolution_mode = pg.convolution_mode
        self.output_conv = pg.output_conv or (in_channels != out_channels)
        self.extent = 2 * pg.KP_extent * radius / config.density_parameter
        k_points = create_kernel_points(1.5 * self.extent, self.num_kernel_points, num_kernels=1, dimension=3,
                                        fixed=pg.fixed_kernel_points).reshape((self.num_kernel_points, 3))
        self.register_buffer('K_points', torch.from_numpy(k_points).type(torch.float32))
        self.grouper = MaskedQueryAndGroup(radius, nsample, use_xyz=False, ret_grouped_xyz=True, normalize_xyz=False)
        self.kernel_weights = weight_variable([self.num_kernel_points, in_channels])
        block = _output_block(in_channels, out_channels, config.bn_momentum, self.output_conv)
        if self.output_conv:
            self.out_conv = block
        else:
            self.out_transform = block

    def forward(self, query_xyz, support_xyz, query_mask, support_mask, support_features):
        if self.KP_influence not in ('constant', 'linear', 'gaussian'):
            raise ValueError('Unknown influence function type (config.KP_influence)')
        if self.convolution_mode != 'sum':
            raise NotImplementedError(f"convolution_mode:{self.convolution_mode} not support in PseudoGrid")
        nbr = self.grouper.neighbors(query_xyz, support_xyz, query_mask, support_mask)
        precision = 1 if runtime.pseudo_grid_precision == 'bf16' else 0
        out = PseudoGridFunction.apply(support_features, self.kernel_weights, query_xyz, support_xyz, query_mask, nbr,
                                       self.K_points, self.extent, self.KP_influence, precision)
        return self.out_conv(out) if self.output_conv else self.out_transform(out)


class LocalAggregation(nn.Module):
    def __init__(self, in_channels, out_channels, radius, nsample, config):
        super().__init__()
        kind = config.local_aggregation_type
        if kind == 'pospool':
            self.local_aggregation_operator = PosPool(in_channels, out_channels, radius, nsample, config)
        elif kind == 'pseudo_grid':
            self.local_aggregation_operator = PseudoGrid(in_channels, out_channels, radius, nsample, config)
        elif kind in ('adaptive_weight', 'pointwisemlp', 'attention'):
            raise NotImplementedError(
                f'LocalAggregation {kind}: outside the B200 hot path (PosPool / PseudoGrid); use the reference '
                f'operator on top of deep3dpointclouddenoising_b200.pt_custom_ops.pt_utils.MaskedQueryAndGroup')
        else:
            raise NotImplementedError(f'LocalAggregation {kind} not implemented')

    def forward(self, query_xyz, support_xyz, query_mask, support_mask, support_features):
        return self.local_aggregation_operator(query_xyz, support_xyz, query_mask, support_mask, support_features)
