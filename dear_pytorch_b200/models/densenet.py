"""DenseNet-BC (Huang et al. 2017): 121 / 169 / 201.  DenseNet-201 (bs 32) is one of the reference's
benchmark tasks (benchmarks.py:21); it has 604 parameter tensors, the worst case for the reference's
per-parameter kernels."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.fused_bn import FusedBatchNormAct2d


def _bn_relu(c, fused):
    """BatchNorm followed by ReLU: one fused channels-last kernel, or the stock module (+ F.relu)."""
    return FusedBatchNormAct2d(c, relu=True) if fused else nn.BatchNorm2d(c)


class _DenseLayer(nn.Module):
    def __init__(self, cin, growth, bn_size, fused_bn=False):
        super().__init__()
        self.fused = fused_bn
        self.norm1 = _bn_relu(cin, fused_bn)
        self.conv1 = nn.Conv2d(cin, bn_size * growth, 1, bias=False)
        self.norm2 = _bn_relu(bn_size * growth, fused_bn)
        self.conv2 = nn.Conv2d(bn_size * growth, growth, 3, padding=1, bias=False)

    def forward(self, feats):
        x = torch.cat(feats, 1)
        if self.fused:
            return self.conv2(self.norm2(self.conv1(self.norm1(x))))
        x = self.conv1(F.relu(self.norm1(x), inplace=True))
        return self.conv2(F.relu(self.norm2(x), inplace=True))


class _DenseBlock(nn.Module):
    def __init__(self, n, cin, growth, bn_size, fused_bn=False):
        super().__init__()
        self.layers = nn.ModuleList(_DenseLayer(cin + i * growth, growth, bn_size, fused_bn) for i in range(n))

    def forward(self, x):
        feats = [x]
        for layer in self.layers:
            feats.append(layer(feats))
        return torch.cat(feats, 1)


class _Transition(nn.Sequential):
    def __init__(self, cin, cout, fused_bn=False):
        head = [FusedBatchNormAct2d(cin, relu=True), nn.Identity()] if fused_bn else \
            [nn.BatchNorm2d(cin), nn.ReLU(inplace=True)]      # (Identity keeps the state-dict indices)
        super().__init__(*head, nn.Conv2d(cin, cout, 1, bias=False), nn.AvgPool2d(2, 2))


class DenseNet(nn.Module):
    def __init__(self, growth=32, blocks=(6, 12, 24, 16), init_features=64, bn_size=4, num_classes=1000, fused_bn=False):
        super().__init__()
        self.fused = fused_bn
        stem_bn = [FusedBatchNormAct2d(init_features, relu=True), nn.Identity()] if fused_bn else \
            [nn.BatchNorm2d(init_features), nn.ReLU(inplace=True)]
        self.stem = nn.Sequential(nn.Conv2d(3, init_features, 7, stride=2, padding=3, bias=False), *stem_bn,
                                  nn.MaxPool2d(3, 2, 1))
        stages, c = [], init_features
        for i, n in enumerate(blocks):
            stages.append(_DenseBlock(n, c, growth, bn_size, fused_bn))
            c += n * growth
            if i != len(blocks) - 1:
                stages.append(_Transition(c, c // 2, fused_bn))
                c //= 2
        self.stages = nn.Sequential(*stages)
        self.norm = _bn_relu(c, fused_bn)
        self.classifier = nn.Linear(c, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.zeros_(m.bias)

    def forward(self, x):
        x = self.norm(self.stages(self.stem(x)))
        if not self.fused:
            x = F.relu(x, inplace=True)
        return self.classifier(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))


def densenet121(**kw): return DenseNet(32, (6, 12, 24, 16), 64, **kw)
def densenet169(**kw): return DenseNet(32, (6, 12, 32, 32), 64, **kw)
def densenet201(**kw): return DenseNet(32, (6, 12, 48, 32), 64, **kw)
