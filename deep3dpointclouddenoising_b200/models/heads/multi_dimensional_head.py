"""U-Net decoder head (mirror of u_net_arch/models/heads/multi_dimensional_head.py:16-85): four nearest
upsamplings, skip concatenation, 1x1 conv blocks, and a final Conv1d to `num_classes` output dims
(3 for offset regression).  Module names up0..up3, up_conv0..up_conv3, head are the reference's."""
import torch.nn as nn

from ... import neighbors as _neighbors
from ...pt_custom_ops.pt_utils import MaskedUpsample
from ..blocks import FusedSequential, conv_bn


def _block(cin, cout):
    return conv_bn(cin, cout)


class MultiDimHeadResNet(nn.Module):
    def __init__(self, num_classes, width, base_radius, nsamples, isGAN=False):
        super().__init__()
        self.num_classes, self.base_radius, self.nsamples = num_classes, base_radius, nsamples
        for level in range(4):  # up0 works on the coarsest pair (res5 -> res4)
            setattr(self, f"up{level}", MaskedUpsample(radius=(8 >> level) * base_radius, nsample=nsamples[3 - level],
                                                        mode='nearest'))
        self.up_conv0 = _block(24 * width, 4 * width)
        self.up_conv1 = _block(8 * width, 2 * width)
        self.up_conv2 = _block(4 * width, width)
        self.up_conv3 = _block(2 * width, width // 2)
        self.head = FusedSequential(nn.Conv1d(width // 2, width // 2, kernel_size=1, bias=False),
                                    nn.BatchNorm1d(width // 2), nn.ReLU(inplace=True),
                                    nn.Conv1d(width // 2, num_classes, kernel_size=1, bias=True))

    def forward(self, end_points):
        features = end_points['res5_features']
        for level in range(4):
            fine, coarse = f"res{4 - level}", f"res{5 - level}"
            features = getattr(self, f"up{level}")(end_points[fine + '_xyz'], end_points[coarse + '_xyz'],
                                                   end_points[fine + '_mask'], end_points[coarse + '_mask'], features)
            # torch.cat([features, skip], 1) -> up_conv (ref :36-37); the block consumes the pair without the copy
            features = getattr(self, f"up_conv{level}")([features, end_points[fine + '_features']])
        _neighbors.join()  # side-stream neighbourhood work of this forward (neighbors.prebuild) is complete from here on
        return self.head(features)
