"""ResNet-18/34 as the reference trains them (reference smart_compress/models/pytorch/resnet.py:13-303
describes the shape: a torchvision-style residual network whose stem is a stride-1 3x3 convolution —
the CIFAR adaptation — with the max-pool kept, :164-172).

Benchmark driver, written for this repo.  It lives under ``smart_compress.models.pytorch`` on purpose:
the layer predicate wraps every module type defined under that package
(util/pytorch/quantization.py), so residual blocks and the network itself are hooked as well as their
layers — the same tensor can be compressed up to three times, as in the reference (76 forward calls
per step for ResNet-18, 132 for ResNet-34)."""
import torch
from torch import nn

STAGE_WIDTHS = (64, 128, 256, 512)
DEPTHS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}


def conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=False)


class ResidualBlock(nn.Module):
    """Two 3x3 convolutions with batch norm; a 1x1 projection on the skip path when the shape changes."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = conv3x3(cin, cout, stride)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(cout, cout)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, stride=stride, bias=False),
                                            nn.BatchNorm2d(cout))

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        # the projection runs AFTER the residual branch, as in the reference (pytorch/resnet.py:59-74): the order
        # of the hooked calls is part of the boundary (tests/test_hooks_vs_reference.py)
        skip = x if self.downsample is None else self.downsample(x)
        return self.relu(y + skip)


class ResNet(nn.Module):
    def __init__(self, depths, num_classes=10):
        super().__init__()
        self.conv1 = conv3x3(3, STAGE_WIDTHS[0])  # CIFAR stem: 3x3, stride 1
        self.bn1 = nn.BatchNorm2d(STAGE_WIDTHS[0])
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        stages, cin = [], STAGE_WIDTHS[0]
        for i, (width, depth) in enumerate(zip(STAGE_WIDTHS, depths)):
            blocks = [ResidualBlock(cin, width, 1 if i == 0 else 2)]
            blocks += [ResidualBlock(width, width, 1) for _ in range(depth - 1)]
            stages.append(nn.Sequential(*blocks))
            cin = width
        self.layer1, self.layer2, self.layer3, self.layer4 = stages
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(cin, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def build(name: str, num_classes: int = 10) -> ResNet:
    return ResNet(DEPTHS[name], num_classes)
