from .functional import linear, stem_conv, stem_conv_supported, conv2d_tc, conv_tc_supported, conv_tc_wanted, mse_loss, cross_entropy, layer_norm
from .modules import Linear, Conv2dTC, PointwiseConv2d, Conv3x3, StemConv7x7, LayerNorm, MSELoss, CrossEntropyLoss
from .batchnorm import FusedBatchNormAct2d, MaxPool3x3s2

__all__ = ["linear", "conv2d_tc", "conv_tc_supported", "conv_tc_wanted", "Conv2dTC", "PointwiseConv2d", "Conv3x3", "StemConv7x7", "stem_conv", "stem_conv_supported", "mse_loss", "cross_entropy", "layer_norm", "Linear",
           "LayerNorm", "MSELoss", "CrossEntropyLoss", "FusedBatchNormAct2d", "MaxPool3x3s2"]
