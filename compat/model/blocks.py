"""Upsampling decoder stage used by `model.models.decoder` (out of the GNN hot path; cuDNN territory).

Restated for the script-level drop-in only: a stage is `5x5 convolution without bias -> batch norm -> ReLU` followed by a
bilinear x2 resize with aligned corners.  The parameter names a reference checkpoint carries for such a stage are
`transconv1.0.weight` (convolution) and `transconv1.1.{weight,bias,running_mean,running_var,num_batches_tracked}` (batch
norm), so the convolution and the norm live at indices 0 and 1 of a container called `transconv1`; the resize has no state.
"""
import torch.nn as nn
import torch.nn.functional as F


def _conv_bn_relu(cin, cout, ksize):
    layers = [nn.Conv2d(cin, cout, ksize, stride=1, padding=2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
    return nn.Sequential(*layers)


class TransBlock(nn.Module):
    expansion = 1
    SCALE = 2          # every decoder stage doubles the spatial size (8 -> 16 -> ... -> 256)

    def __init__(self, in_channels, out_channels, stride=2, dropout=0.0, kernel_size=5):
        super().__init__()
        if stride != self.SCALE:
            raise ValueError(f"decoder stages upsample by {self.SCALE}; got stride={stride}")
        self.transconv1 = _conv_bn_relu(in_channels, out_channels, kernel_size)

    def forward(self, x):
        y = self.transconv1(x)
        return F.interpolate(y, scale_factor=float(self.SCALE), mode="bilinear", align_corners=True)
