from .resnet import ResNet
