"""ResNet-50 / ResNet-152 (He et al.) for the BASELINE.json perf configs.  Own definition (not
torchvision's module) so the classifier is a ``b200ddp.ops.Linear`` (tcgen05 GEMM in bf16) and the
parameter order/shapes match torchvision's - bucket layouts quoted in SURVEY §2.4-K4 therefore hold.
Convolutions / BatchNorm go to cuDNN through torch (not named hot ops in the north-star)."""
from __future__ import annotations

import os
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops import Conv3x3, FusedBatchNormAct2d, Linear, MaxPool3x3s2, PointwiseConv2d


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: nn.Module | None = None):
        super().__init__()
        # BatchNorm + ReLU (and, for bn3, the shortcut add) are single fused kernels on channels_last CUDA tensors
        self.conv1 = PointwiseConv2d(inplanes, planes)
        self.bn1 = FusedBatchNormAct2d(planes, relu=True)
        self.conv2 = Conv3x3(planes, planes, stride=stride)
        self.bn2 = FusedBatchNormAct2d(planes, relu=True)
        self.conv3 = PointwiseConv2d(planes, planes * self.expansion)
        self.bn3 = FusedBatchNormAct2d(planes * self.expansion, relu=True)
        self.downsample = downsample
        # opt-in (B200DDP_CONV1X1_TC=1 and B200DDP_CONV_BN_FUSE=1): the 1x1 convolutions' GEMM epilogue hands BatchNorm its
        # statistics, so bn1 / bn3 read their input once instead of twice
        self.fuse_stats = os.environ.get("B200DDP_CONV_BN_FUSE", "0") == "1" and self.conv1.use_tc

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        if self.fuse_stats:
            y, part = self.conv1.forward_with_stats(x)
            out = self.bn1(y, partials=part)
            out = self.bn2(self.conv2(out))
            y, part = self.conv3.forward_with_stats(out)
            return self.bn3(y, residual=identity, partials=part)
        out = self.bn1(self.conv1(x))
        out = self.bn2(self.conv2(out))
        return self.bn3(self.conv3(out), residual=identity)


class ResNet(nn.Module):
    def __init__(self, layers: List[int], num_classes: int = 1000, zero_init_residual: bool = False,
                 stem_pad_to: int | None = None):
        super().__init__()
        # Opt-in (B200DDP_STEM_PAD=8): feed the 7x7 stem an input with zero channels appended.  With C=3 the NHWC
        # operand is not 16-byte aligned and cuDNN falls back to an sm80 TF32 kernel plus two layout-convert
        # kernels (0.33 ms of a 5 ms step, profiles/launches_graph.md); C=8 is eligible for the sm_100 bf16
        # kernels.  The parameter keeps torchvision's [64,3,7,7] shape (checkpoints unchanged); the zero taps
        # are appended on the fly.
        if stem_pad_to is None:
            stem_pad_to = int(os.environ.get("B200DDP_STEM_PAD", "0") or 0)
        self.stem_pad_to = stem_pad_to if stem_pad_to > 3 else 0
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = FusedBatchNormAct2d(64, relu=True)
        self.maxpool = MaxPool3x3s2()
        self.layer1 = self._make_layer(64, layers[0], 1)
        self.layer2 = self._make_layer(128, layers[1], 2)
        self.layer3 = self._make_layer(256, layers[2], 2)
        self.layer4 = self._make_layer(512, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = Linear(512 * Bottleneck.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.zeros_(m.bn3.weight)

    def _make_layer(self, planes: int, blocks: int, stride: int) -> nn.Sequential:
        downsample = None
        if stride != 1 or self.inplanes != planes * Bottleneck.expansion:
            downsample = nn.Sequential(PointwiseConv2d(self.inplanes, planes * Bottleneck.expansion, stride=stride),
                                       FusedBatchNormAct2d(planes * Bottleneck.expansion))
        mods = [Bottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * Bottleneck.expansion
        mods += [Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*mods)

    @property
    def input_channels(self) -> int:
        """Channels the fused input kernel should emit (extra ones zero-filled)."""
        return self.stem_pad_to or 3

    def _stem(self, x):
        if not self.stem_pad_to:
            return self.conv1(x)
        extra = self.stem_pad_to - self.conv1.weight.shape[1]
        if x.shape[1] != self.stem_pad_to:            # caller did not pad (plain 3-channel batch)
            x = F.pad(x, (0, 0, 0, 0, 0, self.stem_pad_to - x.shape[1]))
        w = F.pad(self.conv1.weight, (0, 0, 0, 0, 0, extra))
        return F.conv2d(x, w, None, self.conv1.stride, self.conv1.padding)

    def forward(self, x):
        x = self.maxpool(self.bn1(self._stem(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = torch.flatten(self.avgpool(x), 1)
        return self.fc(x)


def resnet50(num_classes: int = 1000, **kw) -> ResNet:
    return ResNet([3, 4, 6, 3], num_classes, **kw)


def resnet152(num_classes: int = 1000, **kw) -> ResNet:
    return ResNet([3, 8, 36, 3], num_classes, **kw)
