from .masked_l1_loss import MaskedL1Loss
from .masked_chamfer_loss import MaskedAdaptiveL1ChamferLoss, MaskedChamferL1Loss, MaskedChamferLoss, masked_chamfer
