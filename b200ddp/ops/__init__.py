from .functional import linear, conv2d_tc, conv_tc_supported, conv_tc_wanted, mse_loss, cross_entropy, layer_norm
from .modules import Linear, Conv2dTC, PointwiseConv2d, Conv3x3, LayerNorm, MSELoss, CrossEntropyLoss
from .batchnorm import FusedBatchNormAct2d, MaxPool3x3s2

__all__ = ["linear", "conv2d_tc", "conv_tc_supported", "conv_tc_wanted", "Conv2dTC", "PointwiseConv2d", "Conv3x3", "mse_loss", "cross_entropy", "layer_norm", "Linear",
           "LayerNorm", "MSELoss", "CrossEntropyLoss", "FusedBatchNormAct2d", "MaxPool3x3s2"]
