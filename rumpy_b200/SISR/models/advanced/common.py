"""B200-native mirror of the reference's common block library
(/root/reference/rumpy/SISR/models/advanced/common.py): same names, constructor signatures, parameter
registration order and state_dict keys; the arithmetic runs in librumpy_b200.so (no torch fallback).

Whole networks (RCAN / EDSR) do not execute these modules one by one: they hand the complete parameter list
to the native executor (rumpy_b200/engine.py).  The per-module `forward`s below exist so the blocks stay
usable on their own (inference), again through the C ABI.
"""
import math

import torch
from torch import nn

from rumpy_b200 import blocks_native as _bn


class Conv2d(nn.Conv2d):
    """nn.Conv2d parameter holder (identical state_dict / init) whose forward is the sm_100a kernel."""

    def forward(self, x):
        return _bn.conv_forward(self, x)


def default_conv(in_channels, out_channels, kernel_size, bias=True):
    """reference common.py:6-9"""
    return Conv2d(in_channels, out_channels, kernel_size, padding=(kernel_size // 2), bias=bias)


class MeanShift(nn.Conv2d):
    """reference common.py:12-20 (unused by RCAN/EDSR; kept for import compatibility, plain 1x1 weights)."""

    def __init__(self, rgb_range, rgb_mean, rgb_std, sign=-1):
        super(MeanShift, self).__init__(3, 3, kernel_size=1)
        std = torch.Tensor(rgb_std)
        self.weight.data = torch.eye(3).view(3, 3, 1, 1)
        self.weight.data.div_(std.view(3, 1, 1, 1))
        self.bias.data = sign * rgb_range * torch.Tensor(rgb_mean)
        self.bias.data.div_(std)
        self.requires_grad = False


class PixelShuffle(nn.PixelShuffle):
    """Marker module: inside Upsampler the shuffle is folded into the preceding conv's TMA store."""


class Upsampler(nn.Sequential):
    """reference common.py:23-48: [conv(C->4C), PixelShuffle(2)] x log2(scale)  or  conv(C->9C), PixelShuffle(3)."""

    def __init__(self, conv, scale, n_feat, bn=False, act=False, bias=True):
        m = []
        if bn or act:
            raise NotImplementedError('rumpy_b200 Upsampler: bn/act variants are not used by RCAN/EDSR')
        if (scale & (scale - 1)) == 0:
            for _ in range(int(math.log(scale, 2))):
                m.append(conv(n_feat, 4 * n_feat, 3, bias))
                m.append(PixelShuffle(2))
        elif scale == 3:
            m.append(conv(n_feat, 9 * n_feat, 3, bias))
            m.append(PixelShuffle(3))
        else:
            raise NotImplementedError
        super(Upsampler, self).__init__(*m)

    def forward(self, x):
        return _bn.upsampler_forward(self, x)


class ResBlock(nn.Module):
    """reference common.py:51-75: conv-ReLU-conv, .mul(res_scale), += x."""

    def __init__(self, conv, n_feats, kernel_size, bias=True, bn=False, act=nn.ReLU(True), res_scale=1.0):
        super(ResBlock, self).__init__()
        if bn:
            raise NotImplementedError('rumpy_b200 ResBlock: bn variant is not used by EDSR')
        m = []
        for i in range(2):
            m.append(conv(n_feats, n_feats, kernel_size, bias=bias))
            if i == 0:
                m.append(act)
        self.body = nn.Sequential(*m)
        self.res_scale = res_scale

    def forward(self, x):
        return _bn.resblock_forward(self, x)
