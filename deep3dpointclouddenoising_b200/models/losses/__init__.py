from .masked_l1_loss import MaskedL1Loss
