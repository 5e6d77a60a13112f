"""Decoder block of the reference (dgl/model/blocks.py:5-25): 5x5 conv (no bias) + BN + ReLU, then bilinear x2.
Out of the hot path (cuDNN territory); restated here only so `model.models` is a complete, importable module with the
reference's parameter names (`transconv1.0.weight`, `transconv1.1.*`)."""
import torch.nn as nn


class TransBlock(nn.Module):
    expansion = 1

    def __init__(self, in_channels, out_channels, stride=1, dropout=0.0, kernel_size=3):
        super().__init__()
        if stride != 2:
            raise ValueError("the reference only ever builds stride-2 TransBlocks (dgl/model/models.py:61-65)")
        conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=1, padding=2, bias=False)
        self.transconv1 = nn.Sequential(conv, nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))
        self.Upsample = nn.Upsample(scale_factor=(stride, stride), mode='bilinear', align_corners=True)

    def forward(self, x):
        return self.Upsample(self.transconv1(x))
