from .functional import linear, conv1x1, conv3x3, mse_loss, cross_entropy, layer_norm
from .modules import Linear, PointwiseConv2d, Conv3x3, LayerNorm, MSELoss, CrossEntropyLoss
from .batchnorm import FusedBatchNormAct2d, MaxPool3x3s2

__all__ = ["linear", "conv1x1", "conv3x3", "PointwiseConv2d", "Conv3x3", "mse_loss", "cross_entropy", "layer_norm", "Linear", "LayerNorm", "MSELoss", "CrossEntropyLoss", "FusedBatchNormAct2d", "MaxPool3x3s2"]
