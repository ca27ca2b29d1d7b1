"""Inception-v4 (Szegedy et al. 2016, "Inception-v4, Inception-ResNet and the Impact of Residual
Connections on Learning"), 299x299 input, 42.7 M parameters.

The reference vendors the Cadene ``pretrainedmodels`` implementation six times (``*/inceptionv4.py``)
for its ``--model inceptionv4`` task (benchmarks.py:21).  This is an independent implementation from
the paper's block diagrams, expressed with a small branch-table helper.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class ConvBN(nn.Sequential):
    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__(nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=False),
                         nn.BatchNorm2d(cout, eps=0.001, momentum=0.1), nn.ReLU(inplace=True))


def _chain(cin, specs):
    """specs: [(cout, kernel, stride, padding), ...] -> Sequential of ConvBN."""
    layers = []
    for cout, k, s, p in specs:
        layers.append(ConvBN(cin, cout, k, s, p))
        cin = cout
    return nn.Sequential(*layers)


class Branches(nn.Module):
    """Run parallel branches on the same input and concatenate along channels."""

    def __init__(self, *branches):
        super().__init__()
        self.branches = nn.ModuleList(branches)

    def forward(self, x):
        return torch.cat([b(x) for b in self.branches], 1)


class Stem(nn.Module):
    def __init__(self):
        super().__init__()
        self.pre = _chain(3, [(32, 3, 2, 0), (32, 3, 1, 0), (64, 3, 1, 1)])                 # 149 -> 147
        self.mix1 = Branches(nn.MaxPool2d(3, 2), ConvBN(64, 96, 3, 2))                        # 160 @ 73
        self.mix2 = Branches(_chain(160, [(64, 1, 1, 0), (96, 3, 1, 0)]),
                             _chain(160, [(64, 1, 1, 0), (64, (1, 7), 1, (0, 3)), (64, (7, 1), 1, (3, 0)), (96, 3, 1, 0)]))
        self.mix3 = Branches(ConvBN(192, 192, 3, 2), nn.MaxPool2d(3, 2))                       # 384 @ 35

    def forward(self, x):
        return self.mix3(self.mix2(self.mix1(self.pre(x))))


def _avgpool_conv(cin, cout):
    return nn.Sequential(nn.AvgPool2d(3, 1, 1, count_include_pad=False), ConvBN(cin, cout, 1))


def inception_a():
    return Branches(ConvBN(384, 96, 1),
                    _chain(384, [(64, 1, 1, 0), (96, 3, 1, 1)]),
                    _chain(384, [(64, 1, 1, 0), (96, 3, 1, 1), (96, 3, 1, 1)]),
                    _avgpool_conv(384, 96))


def reduction_a():
    return Branches(ConvBN(384, 384, 3, 2),
                    _chain(384, [(192, 1, 1, 0), (224, 3, 1, 1), (256, 3, 2, 0)]),
                    nn.MaxPool2d(3, 2))                                                        # 1024 @ 17


def inception_b():
    return Branches(ConvBN(1024, 384, 1),
                    _chain(1024, [(192, 1, 1, 0), (224, (1, 7), 1, (0, 3)), (256, (7, 1), 1, (3, 0))]),
                    _chain(1024, [(192, 1, 1, 0), (192, (7, 1), 1, (3, 0)), (224, (1, 7), 1, (0, 3)),
                                  (224, (7, 1), 1, (3, 0)), (256, (1, 7), 1, (0, 3))]),
                    _avgpool_conv(1024, 128))


def reduction_b():
    return Branches(_chain(1024, [(192, 1, 1, 0), (192, 3, 2, 0)]),
                    _chain(1024, [(256, 1, 1, 0), (256, (1, 7), 1, (0, 3)), (320, (7, 1), 1, (3, 0)), (320, 3, 2, 0)]),
                    nn.MaxPool2d(3, 2))                                                        # 1536 @ 8


class _SplitTail(nn.Module):
    """A trunk followed by two sibling convolutions whose outputs are concatenated."""

    def __init__(self, trunk, cin, cout):
        super().__init__()
        self.trunk = trunk
        self.a = ConvBN(cin, cout, (1, 3), 1, (0, 1))
        self.b = ConvBN(cin, cout, (3, 1), 1, (1, 0))

    def forward(self, x):
        t = self.trunk(x)
        return torch.cat([self.a(t), self.b(t)], 1)


def inception_c():
    return Branches(ConvBN(1536, 256, 1),
                    _SplitTail(ConvBN(1536, 384, 1), 384, 256),
                    _SplitTail(_chain(1536, [(384, 1, 1, 0), (448, (3, 1), 1, (1, 0)), (512, (1, 3), 1, (0, 1))]), 512, 256),
                    _avgpool_conv(1536, 256))


class InceptionV4(nn.Module):
    def __init__(self, num_classes=1000, dropout=0.2):
        super().__init__()
        blocks = [Stem()] + [inception_a() for _ in range(4)] + [reduction_a()] + \
                 [inception_b() for _ in range(7)] + [reduction_b()] + [inception_c() for _ in range(3)]
        self.features = nn.Sequential(*blocks)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.dropout = nn.Dropout(dropout)
        self.last_linear = nn.Linear(1536, num_classes)

    def forward(self, x):
        x = self.pool(self.features(x))
        return self.last_linear(self.dropout(torch.flatten(x, 1)))


def inceptionv4(num_classes=1000, **kw):
    return InceptionV4(num_classes=num_classes, **kw)
