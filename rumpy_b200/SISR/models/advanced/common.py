"""B200-native mirror of the reference's common block library
(/root/reference/rumpy/SISR/models/advanced/common.py): same names, constructor signatures, parameter
registration order and state_dict keys; the arithmetic runs in librumpy_b200.so (no torch fallback).

Whole networks (RCAN / EDSR) do not execute these modules one by one: they hand the complete parameter list
to the native executor (rumpy_b200/engine.py).  The per-module `forward`s below exist so the blocks stay
usable on their own (inference), again through the C ABI.
"""
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
    """reference common.py:12-20: fixed 1x1 conv `(x + sign * range * mean) / std` (unused by RCAN / EDSR; kept for
    import compatibility; the parameters are plain 1x1 weights, frozen)."""

    def __init__(self, rgb_range, rgb_mean, rgb_std, sign=-1):
        super().__init__(3, 3, kernel_size=1)
        inv_std = 1.0 / torch.as_tensor(rgb_std, dtype=torch.float32)
        mean = torch.as_tensor(rgb_mean, dtype=torch.float32)
        with torch.no_grad():
            self.weight.copy_(torch.diag(inv_std).reshape(3, 3, 1, 1))
            self.bias.copy_(sign * rgb_range * mean * inv_std)
        self.requires_grad = False


class PixelShuffle(nn.PixelShuffle):
    """Marker module: inside Upsampler the shuffle is folded into the preceding conv's TMA store."""


def _shuffle_factors(scale):
    """Upsampling stages of the reference's Upsampler (common.py:27-45): 2, 2, ... for powers of two, one x3 stage."""
    if scale >= 1 and scale & (scale - 1) == 0:
        return [2] * (scale.bit_length() - 1)
    if scale == 3:
        return [3]
    raise NotImplementedError(f'Upsampler: scale {scale} (powers of two and 3 only, like the reference)')


class Upsampler(nn.Sequential):
    """reference common.py:23-48: per stage r a `conv(C -> r*r*C)` followed by `PixelShuffle(r)`; the Sequential
    indices (0, 2 hold parameters at x4) are part of the state_dict contract."""

    def __init__(self, conv, scale, n_feat, bn=False, act=False, bias=True):
        if bn or act:
            raise NotImplementedError('rumpy_b200 Upsampler: bn/act variants are not used by RCAN/EDSR')
        stages = []
        for r in _shuffle_factors(int(scale)):
            stages += [conv(n_feat, r * r * n_feat, 3, bias), PixelShuffle(r)]
        super().__init__(*stages)

    def forward(self, x):
        return _bn.upsampler_forward(self, x)


class ResBlock(nn.Module):
    """reference common.py:51-75: conv-ReLU-conv, .mul(res_scale), += x (`body.0`, `body.2` hold the convs)."""

    def __init__(self, conv, n_feats, kernel_size, bias=True, bn=False, act=nn.ReLU(True), res_scale=1.0):
        super().__init__()
        if bn:
            raise NotImplementedError('rumpy_b200 ResBlock: bn variant is not used by EDSR')
        self.body = nn.Sequential(conv(n_feats, n_feats, kernel_size, bias=bias), act,
                                  conv(n_feats, n_feats, kernel_size, bias=bias))
        self.res_scale = res_scale

    def forward(self, x):
        return _bn.resblock_forward(self, x)
