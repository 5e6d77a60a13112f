"""Replacement for the reference's dgl/model/models.py: same classes / constructor arguments / state_dict keys, with the
GNN slice (models.py:180-189) on librobotgnn.  Encoder and decoder are plain torchvision / cuDNN modules, out of the hot
path, restated with the reference's attribute names so whole-module checkpoints (training.py:345-355) stay loadable."""
import copy

import torch
import torch.nn as nn
from torchvision import models as tvm

import robot_gnn_b200 as rg

from .blocks import TransBlock

edge_encoder = rg.edge_encoder          # dgl/model/models.py:142-155
GCN = rg.GCN                            # dgl/model/models.py:213-226
GraphConv = rg.GraphConv                # imported, unused by the reference (models.py:7)

_MOBILENET_STAGES = ((0, 1), (2, 3), (4, 5, 6), (7, 8, 9, 10, 11, 12, 13), (14, 15, 16, 17, 18))   # models.py:18


class encoder(nn.Module):
    """Backbone wrapper (models.py:11-50): five stages -> list of feature maps, the last one [N, F, S/32, S/32]."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        stages = nn.ModuleList(nn.ModuleList() for _ in range(5))
        if opt.backbone == 'mobilenetv2':
            net = tvm.mobilenet_v2(pretrained=opt.pretrained)
            for si, idxs in enumerate(_MOBILENET_STAGES):
                for i in idxs:
                    stages[si].append(net.features[i])
        elif opt.backbone in ('resnet50', 'resnet18'):
            assert opt.skip_level is False or opt.skip_level == False  # noqa: E712  (models.py:24,31)
            net = getattr(tvm, opt.backbone)(pretrained=opt.pretrained)
            net.fc, net.avgpool = nn.Sequential(), nn.Sequential()
            stages[0].append(net)
        else:
            raise AssertionError("unknown backbone")           # models.py:37
        self.image_encoder = stages

    def forward(self, images):
        s = self.opt.image_size
        x = images.view(-1, 3, s, s)
        outs = []
        resnet = self.opt.backbone in ('resnet18', 'resnet50')
        for stage in self.image_encoder:
            for layer in stage:
                x = layer(x)
            if resnet:
                x = x.view(x.size(0), -1, s // 32, s // 32)
            outs.append(x)
        return outs


class decoder(nn.Module):
    """Five TransBlocks + 1x1 head (models.py:53-83)."""

    def __init__(self, opt, nfeature_image):
        super().__init__()
        self.opt = opt
        self.nfeature_image = nfeature_image
        self.nfeature_array = [nfeature_image // (1 << i) for i in range(6)]
        blocks = [TransBlock(self.nfeature_array[i], self.nfeature_array[i + 1], 2, kernel_size=5) for i in range(5)]
        blocks.append(nn.Conv2d(self.nfeature_array[5], opt.output_dim, kernel_size=1))
        self.image_decoder = nn.ModuleList(blocks)
        if opt.skip_level:
            self.add_conv = nn.ModuleList(nn.Conv2d(c, self.nfeature_array[i + 1], kernel_size=1)
                                          for i, c in enumerate((96, 32, 24, 16)))

    def forward(self, h, h_list=None):
        for i, blk in enumerate(self.image_decoder):
            h = blk(h)
            if i <= 3 and h_list is not None and self.opt.skip_level:
                h = h + self.add_conv[i](h_list[-i - 2])
        return h.view(-1, self.opt.camera_num, self.opt.output_dim, self.opt.image_size, self.opt.image_size)


class single_view_model(nn.Module):
    """models.py:86-118 (no GNN; kept so `-model single_view` keeps working from the same module)."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.encoder = encoder(opt)
        if opt.task == 'depthseg':
            dopt = copy.copy(opt)
            dopt.output_dim = 1
            self.depth_decoder = decoder(dopt, opt.feature_dim)
            self.seg_decoder = decoder(opt, opt.feature_dim)
        else:
            self.decoder = decoder(opt, opt.feature_dim)

    def forward(self, image):
        h_list = self.encoder(image)
        skip = h_list if self.opt.skip_level else None
        if self.opt.task == 'depthseg':
            return self.depth_decoder(h_list[-1], skip), self.seg_decoder(h_list[-1], skip)
        return self.decoder(h_list[-1], skip)


class multi_view_dgl_model(nn.Module):
    """models.py:157-205.  Parameters keep the reference's flat names (gcn1.*, conv1.*, gcn2.*, conv2.*).
    return_key='images' (default) is the intended aggregate; 'image' reproduces the literal reference, whose GCN.forward
    returns its input (models.py:226)."""

    def __init__(self, opt, return_key="images"):
        super().__init__()
        self.opt = opt
        f = opt.feature_dim
        self.encoder = encoder(opt)
        self.gcn1 = GCN(opt, return_key)
        if opt.compress_gcn:
            self.conv1 = nn.Conv2d(f * 2, f, kernel_size=1)
            self.decoder = decoder(opt, f)
        else:
            self.decoder = decoder(opt, f * 2)
        if opt.multi_gcn:
            assert opt.compress_gcn                                          # models.py:169
            self.gcn2 = GCN(opt, return_key)
            self.conv2 = nn.Conv2d(f * 2, f, kernel_size=1)

    def forward(self, g):
        with g.local_scope():
            s = self.opt.image_size
            image = g.ndata['image'].view(-1, self.opt.camera_num, 3, s, s)
            h_list = self.encoder(image)
            h = h_list[-1]
            h = h.view(-1, h.size(-3), h.size(-2), h.size(-1))
            # models.py:180-189; under torch.no_grad() (dgl/eval.py) small batches replay a captured CUDA graph of the slice
            h = rg.gnn_slice_auto(self, g, h, compute_dtype=getattr(self, 'gnn_dtype', None))   # channels-last / other-dtype maps pass the seams
            if self.opt.task == 'depthseg':
                # the reference never builds depth_decoder / seg_decoder for this model (models.py:191-193)
                raise AttributeError("multi_view_dgl_model has no depth_decoder/seg_decoder (reference defect, task=depthseg)")
            return self.decoder(h, h_list if self.opt.skip_level else None)
