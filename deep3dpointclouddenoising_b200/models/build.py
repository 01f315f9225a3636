"""Model factory for offset regression (mirror of u_net_arch/models/build.py:42-67, 236-262)."""
import torch.nn as nn

from .backbones import ResNet
from .heads import MultiDimHeadResNet
from .losses import MaskedAdaptiveL1ChamferLoss, MaskedChamferL1Loss, MaskedChamferLoss, MaskedL1Loss

OFFSET_REG_DIM = 3


class OffsetRegressionModel(nn.Module):
    def __init__(self, config, backbone, head, num_classes, input_features_dim, radius, sampleDl, nsamples, npoints,
                 width=144, depth=2, bottleneck_ratio=2):
        super().__init__()
        if input_features_dim == 0:
            input_features_dim = 3  # the dataset feeds xyz as features (offset_dataset.py:726)
        if backbone != 'resnet':
            raise NotImplementedError(f"Backbone {backbone} not implemented in Offset Regression Model")
        self.backbone = ResNet(config, input_features_dim, radius, sampleDl, nsamples, npoints, width=width,
                               depth=depth, bottleneck_ratio=bottleneck_ratio)
        if head != 'offset_reg_head':
            raise NotImplementedError(f"Head {backbone} not implemented in Offset Regression Model")
        # attribute name kept from the reference (state-dict prefix 'segmentation_head.')
        self.segmentation_head = MultiDimHeadResNet(OFFSET_REG_DIM, width, radius, nsamples, isGAN=config.GAN)

    def forward(self, xyz, mask, features):
        return self.segmentation_head(self.backbone(xyz, mask, features))

    def prefetch_neighbors(self, xyz, mask, with_csr=None):
        """Training loops call this between forward and backward with the NEXT batch's (xyz, mask): its neighbourhood
        pyramid is built on a side stream while backward runs, and the next forward on those tensors adopts it."""
        self.backbone.prefetch_neighbors(xyz, mask, with_csr)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, (nn.Conv1d, nn.Conv2d)):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)


def build_offset_regression(config):
    model = OffsetRegressionModel(config, config.backbone, config.head, config.num_classes, config.input_features_dim,
                                  config.radius, config.sampleDl, config.nsamples, config.npoints, config.width,
                                  config.depth, config.bottleneck_ratio)
    if config.loss == 'L1':
        criterion = MaskedL1Loss()
    elif config.loss is None:
        raise ValueError("Please specify a loss in the config file")
    elif config.loss == 'chamfer_L1':  # build.py:51-62: forward(pred, target, mask, points) for all of these
        criterion = MaskedChamferL1Loss()
    elif config.loss == 'chamfer':
        criterion = MaskedChamferLoss()
    elif config.loss == 'chamfer_sparse':
        criterion = MaskedChamferLoss(norm_type='L1')
    elif config.loss == 'l1_chamfer_sparse':
        criterion = MaskedChamferL1Loss(norm_type='L1')
    elif config.loss == 'l1_chamfer_adaptive_to_chamfer':
        criterion = MaskedAdaptiveL1ChamferLoss(converging_to='chamfer')
    elif config.loss == 'l1_chamfer_adaptive_to_l1':
        criterion = MaskedAdaptiveL1ChamferLoss(converging_to='L1')
    else:
        raise ValueError(f"The loss {config.loss} is not implemented")
    return model, criterion
