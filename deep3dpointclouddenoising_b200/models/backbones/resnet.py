"""Point ResNet backbone (mirror of u_net_arch/models/backbones/resnet.py: Bottleneck :22-68, ResNet :71-188).

Same module tree and parameter names (conv1, la1, btnk1, layer{1..4}.strided_bottleneck,
layer{1..4}.bottlneck{i} — the reference's spelling — with conv1 / local_aggregation / conv2 / shortcut /
maxpool inside each bottleneck) and the same end_points dictionary, so state dicts are interchangeable.
The 1x1 convolutions and BatchNorm stay in PyTorch (outside the hot path); everything
between them — subsampling, neighbour lists, max-pool, local aggregation — runs on the fused kernels.
One forward builds each distinct neighbour list once (neighbors.cache) instead of 14 ball queries.
"""
import torch
import torch.nn as nn

from ... import neighbors as _neighbors
from ...fused import stage_marker
from ...utils.config import runtime
from ...pt_custom_ops.pt_utils import MaskedMaxPool
from ..blocks import conv_bn
from ..local_aggregation_operators import LocalAggregation


def _conv_bn(cin, cout, momentum, relu):
    return conv_bn(cin, cout, momentum, relu)


class MultiInputSequential(nn.Sequential):
    def forward(self, *inputs):
        for module in self._modules.values():
            inputs = module(*inputs)
        return inputs


class Bottleneck(nn.Module):
    def __init__(self, in_channels, out_channels, bottleneck_ratio, radius, nsample, config, downsample=False,
                 sampleDl=None, npoint=None):
        super().__init__()
        self.in_channels, self.out_channels, self.downsample = in_channels, out_channels, downsample
        mid = out_channels // bottleneck_ratio
        if downsample:
            self.maxpool = MaskedMaxPool(npoint, radius, nsample, sampleDl)
        self.conv1 = _conv_bn(in_channels, mid, config.bn_momentum, relu=True)
        self.local_aggregation = LocalAggregation(mid, mid, radius, nsample, config)
        self.conv2 = _conv_bn(mid, out_channels, config.bn_momentum, relu=False)
        self.relu = nn.ReLU(inplace=True)
        if in_channels != out_channels:
            self.shortcut = _conv_bn(in_channels, out_channels, config.bn_momentum, relu=False)

    def forward(self, xyz, mask, features):
        if self.downsample:
            query_xyz, query_mask, identity = self.maxpool(xyz, mask, features)
        else:
            query_xyz, query_mask, identity = xyz, mask, features
        out = self.conv1(features)
        out = self.local_aggregation(query_xyz, xyz, query_mask, mask, out)
        if self.in_channels != self.out_channels:
            identity = self.shortcut(identity)
        # conv2 -> BN, + identity, ReLU (resnet.py:58-66): the add and the ReLU ride on the BN apply pass
        return query_xyz, query_mask, self.conv2(out, residual=identity, final_relu=True)


class ResNet(nn.Module):
    def __init__(self, config, input_features_dim, radius, sampleDl, nsamples, npoints, width=144, depth=2,
                 bottleneck_ratio=2):
        super().__init__()
        self.input_features_dim = input_features_dim
        half = width // 2
        self.conv1 = _conv_bn(input_features_dim, half, config.bn_momentum, relu=True)
        self.la1 = LocalAggregation(half, half, radius, nsamples[0], config)
        self.btnk1 = Bottleneck(half, width, bottleneck_ratio, radius, nsamples[0], config)
        # what one forward will ask neighbors.py for, with the very values handed to the modules below (cache keys)
        self._radius0, self._nsample0, self._stages = radius, nsamples[0], []
        # the staged-tile PosPool kernels need a processing order per level (fused.PosPoolFunction)
        self._with_order = (config.local_aggregation_type == 'pospool' and config.pospool.position_embedding == 'xyz'
                            and config.pospool.reduction in ('avg', 'mean'))
        # four strided stages: each halves the resolution (grid cell x2) and doubles radius and width
        for stage in range(4):
            sampleDl *= 2
            self._stages.append((sampleDl, npoints[stage], radius, nsamples[stage], radius * 2, nsamples[stage + 1]))
            layer = MultiInputSequential()
            layer.add_module("strided_bottleneck",
                             Bottleneck(width, 2 * width, bottleneck_ratio, radius, nsamples[stage], config,
                                        downsample=True, sampleDl=sampleDl, npoint=npoints[stage]))
            radius *= 2
            width *= 2
            for i in range(depth - 1):
                layer.add_module(f"bottlneck{i}",
                                 Bottleneck(width, width, bottleneck_ratio, radius, nsamples[stage + 1], config))
            setattr(self, f"layer{stage + 1}", layer)

    def forward(self, xyz, mask, features, end_points=None):
        # a pyramid prefetched for exactly these tensors during the previous step (prefetch_neighbors) is adopted as is
        if not _neighbors.adopt(xyz, mask):
            _neighbors.cache.clear()  # neighbour lists are valid for one forward only
            if runtime.prefetch_neighbors and xyz.is_cuda:
                # the whole pyramid (+ inverse maps when gradients are on) goes to a side stream and overlaps with the
                # convolutions / BatchNorm / aggregations below; consumers wait on per-item events (neighbors.py)
                _neighbors.prebuild(xyz, mask, self._radius0, self._nsample0, self._stages, torch.is_grad_enabled(),
                                    self._with_order)
        if not end_points:
            end_points = {}
        try:
            return self._forward(xyz, mask, features, end_points)
        finally:
            # the side-stream work of this forward is joined here — also when the backbone is used without the head and
            # when the forward raises — so the stream is never left forked (a later graph capture would fail).  The
            # pyramid is long finished by the time the last stage has run.
            _neighbors.join()

    def prefetch_neighbors(self, xyz, mask, with_csr=None):
        """Builds the pyramid of the NEXT batch now, on the side stream (training: call it between this step's forward and
        its backward; inference: before this batch's forward): the next forward on these very tensors adopts it instead
        of building its own (neighbors.prefetch).  with_csr: also the inverse maps backward needs (default: grad mode)."""
        with_csr = torch.is_grad_enabled() if with_csr is None else with_csr
        _neighbors.prefetch(xyz, mask, self._radius0, self._nsample0, self._stages, with_csr, self._with_order)

    def _forward(self, xyz, mask, features, end_points):
        features = self.conv1(features)
        features = self.la1(xyz, xyz, mask, mask, features)
        xyz, mask, features = self.btnk1(xyz, mask, features)
        end_points['res1_xyz'], end_points['res1_mask'], end_points['res1_features'] = xyz, mask, features
        for stage in range(4):
            # backward leaves layer{stage+1} when the gradient passes this marker: with data parallelism the gradients of
            # the stage (and of everything after it) go to the all-reduce then (distributed.overlap_with_backward)
            xyz, mask, features = getattr(self, f"layer{stage + 1}")(xyz, mask, stage_marker(features, f"layer{stage + 1}"))
            tag = f"res{stage + 2}"
            end_points[tag + '_xyz'], end_points[tag + '_mask'], end_points[tag + '_features'] = xyz, mask, features
        # the head's first op consumes res5: its backward is the head's last
        end_points['res5_features'] = stage_marker(end_points['res5_features'], "head")
        return end_points
