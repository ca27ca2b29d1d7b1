"""VGG (Simonyan & Zisserman 2014), configurations A/D/E without batch-norm.

VGG-16 is BASELINE.json's communication-bound config: its first fully connected layer alone is a
392 MB fp32 fusion bucket (102.76 M parameters), the bandwidth test of Kernel A / Kernel B.
"""
import torch
import torch.nn as nn

_CFG = {
    "A": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "D": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
    "E": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"],
}


class VGG(nn.Module):
    def __init__(self, cfg, num_classes=1000, dropout=0.5):
        super().__init__()
        layers, cin = [], 3
        for v in cfg:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=True)]
                cin = v
        self.features = nn.Sequential(*layers)
        self.avgpool = nn.AdaptiveAvgPool2d((7, 7))
        self.classifier = nn.Sequential(
            nn.Linear(512 * 7 * 7, 4096), nn.ReLU(True), nn.Dropout(dropout),
            nn.Linear(4096, 4096), nn.ReLU(True), nn.Dropout(dropout),
            nn.Linear(4096, num_classes))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        x = self.avgpool(self.features(x))
        return self.classifier(torch.flatten(x, 1))


def vgg11(**kw): return VGG(_CFG["A"], **kw)
def vgg16(**kw): return VGG(_CFG["D"], **kw)
def vgg19(**kw): return VGG(_CFG["E"], **kw)
