"""ResNet family (He et al. 2015), v1.5 bottleneck (stride on the 3x3 conv).

The reference benchmarks ``torchvision.models.resnet50`` (dear/imagenet_benchmark.py:78-82);
this is an independent implementation of the same architecture (25,557,032 parameters for
ResNet-50) so the benchmark does not depend on torchvision.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..ops.fused_bn import FusedBatchNormAct2d


def _bn(c, relu, fused):
    """BatchNorm (+ReLU when ``relu``): the fused channels-last kernel or the stock modules."""
    return FusedBatchNormAct2d(c, relu=relu) if fused else nn.BatchNorm2d(c)


def _conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)


def _conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 1, stride=stride, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, width, stride=1, downsample=None, fused_bn=False):
        super().__init__()
        self.fused = fused_bn
        self.conv1 = _conv3x3(cin, width, stride)
        self.bn1 = _bn(width, True, fused_bn)
        self.conv2 = _conv3x3(width, width)
        self.bn2 = _bn(width, True, fused_bn)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        if self.fused:
            out = self.bn1(self.conv1(x))                       # BN + ReLU in one kernel
            return self.bn2(self.conv2(out), residual=idt)      # BN + add + ReLU in one kernel
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + idt)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, width, stride=1, downsample=None, fused_bn=False):
        super().__init__()
        self.fused = fused_bn
        self.conv1 = _conv1x1(cin, width)
        self.bn1 = _bn(width, True, fused_bn)
        self.conv2 = _conv3x3(width, width, stride)
        self.bn2 = _bn(width, True, fused_bn)
        self.conv3 = _conv1x1(width, width * 4)
        self.bn3 = _bn(width * 4, True, fused_bn)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        if self.fused:
            out = self.bn1(self.conv1(x))
            out = self.bn2(self.conv2(out))
            return self.bn3(self.conv3(out), residual=idt)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + idt)


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=1000, zero_init_residual=False, fused_bn=False):
        super().__init__()
        self.inplanes = 64
        self.fused_bn = fused_bn
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = _bn(64, True, fused_bn)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)
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
                elif isinstance(m, BasicBlock):
                    nn.init.zeros_(m.bn2.weight)

    def _make_layer(self, block, width, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != width * block.expansion:
            downsample = nn.Sequential(_conv1x1(self.inplanes, width * block.expansion, stride),
                                       _bn(width * block.expansion, False, self.fused_bn))
        layers = [block(self.inplanes, width, stride, downsample, fused_bn=self.fused_bn)]
        self.inplanes = width * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, width, fused_bn=self.fused_bn))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.bn1(self.conv1(x))
        x = self.maxpool(x if self.fused_bn else self.relu(x))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def resnet18(**kw): return ResNet(BasicBlock, [2, 2, 2, 2], **kw)
def resnet34(**kw): return ResNet(BasicBlock, [3, 4, 6, 3], **kw)
def resnet50(**kw): return ResNet(Bottleneck, [3, 4, 6, 3], **kw)
def resnet101(**kw): return ResNet(Bottleneck, [3, 4, 23, 3], **kw)
def resnet152(**kw): return ResNet(Bottleneck, [3, 8, 36, 3], **kw)
