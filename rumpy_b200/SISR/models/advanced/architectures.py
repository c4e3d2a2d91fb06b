"""B200-native mirror of the reference's RCAN / EDSR architectures
(/root/reference/rumpy/SISR/models/advanced/architectures.py:24-257).

Same class names, constructor signatures, module tree, parameter registration order and state_dict keys as the
reference (a reference checkpoint loads with strict=True and vice versa), same forward contract
(fp32 NCHW in -> fp32 NCHW out).  The arithmetic is NOT torch: `RCAN.forward` / `EDSR.forward` hand the
parameter list to the native whole-network executor (rumpy_b200/engine.py -> librumpy_b200.so), which runs the
tcgen05 implicit-GEMM convolutions, the fused channel attention and the pixel-shuffle stores on sm_100a.
"""
from collections import OrderedDict

import torch
from torch import nn

from rumpy_b200 import blocks_native as _bn
from rumpy_b200 import engine as _engine
from rumpy_b200.SISR.models.advanced import common
from rumpy_b200.SISR.models.advanced.HAN_blocks import CSAM_Module, LAM_Module
from rumpy_b200.trunk_function import trunk_apply


class CALayer(nn.Module):
    """Channel attention (reference architectures.py:24-44)."""

    def __init__(self, channel, reduction=16):
        super(CALayer, self).__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)       # parameter-less; kept so the module tree matches
        self.conv_du = nn.Sequential(
            nn.Conv2d(channel, channel // reduction, 1, padding=0, bias=True),
            nn.ReLU(inplace=True),
            nn.Conv2d(channel // reduction, channel, 1, padding=0, bias=True),
            nn.Sigmoid()
        )

    def forward(self, x):
        return _bn.ca_forward(self, x)


class RCAB(nn.Module):
    """Residual channel attention block (reference architectures.py:60-84).  `res_scale` is stored and, as in
    the reference (:79, :81-84), never applied."""

    def __init__(self, conv, n_feat, kernel_size, reduction, bias=True, bn=False, act=nn.ReLU(True), res_scale=1):
        super(RCAB, self).__init__()
        if bn:
            raise NotImplementedError('rumpy_b200 RCAB: bn variant is not used by RCAN')
        modules_body = []
        for i in range(2):
            modules_body.append(conv(n_feat, n_feat, kernel_size, bias=bias))
            if i == 0:
                modules_body.append(act)
        modules_body.append(CALayer(n_feat, reduction))
        self.body = nn.Sequential(*modules_body)
        self.res_scale = res_scale

    def forward(self, x):
        return _bn.rcab_forward(self, x)


class ResidualGroup(nn.Module):
    """reference architectures.py:107-124"""

    def __init__(self, conv, n_feat, kernel_size, reduction, act, res_scale, n_resblocks):
        super(ResidualGroup, self).__init__()
        modules_body = [
            RCAB(conv, n_feat, kernel_size, reduction, bias=True, bn=False, act=act, res_scale=res_scale)
            for _ in range(n_resblocks)]
        modules_body.append(conv(n_feat, n_feat, kernel_size))
        self.body = nn.Sequential(*modules_body)

    def forward(self, x):
        return _bn.resgroup_forward(self, x)


class _NativeTrunk(nn.Module):
    """Shared machinery: lazily builds the native engine for the module's parameter list."""

    _engine_obj = None

    def _engine_kwargs(self):
        raise NotImplementedError

    def native_engine(self):
        eng = self._engine_obj
        if eng is None:
            params = list(self.parameters())
            arch, kw = self._engine_kwargs()
            eng = _engine.TrunkEngine(arch, params, **kw)
            object.__setattr__(self, '_engine_obj', eng)   # not a submodule / not in state_dict
        return eng

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() re-create parameter storage: the cached engine (flat views, packed weights,
        # launch plans) is rebuilt on next use
        object.__setattr__(self, '_engine_obj', None)
        return super()._apply(fn, *args, **kwargs)

    def forward(self, x):
        return trunk_apply(self, x)

    def reset_parameters(self):
        pass


class RCAN(_NativeTrunk):
    """reference architectures.py:140-195"""

    def __init__(self, n_resblocks=20, n_resgroups=10, n_feats=64, in_feats=3, out_feats=3, scale=4, reduction=16,
                 res_scale=1.0, **kwargs):
        super(RCAN, self).__init__()
        kernel_size = 3
        act = nn.ReLU(True)
        modules_head = [common.default_conv(in_feats, n_feats, kernel_size)]
        modules_body = [
            ResidualGroup(common.default_conv, n_feats, kernel_size, reduction, act=act, res_scale=res_scale,
                          n_resblocks=n_resblocks) for _ in range(n_resgroups)]
        modules_body.append(common.default_conv(n_feats, n_feats, kernel_size))
        modules_tail = [
            common.Upsampler(common.default_conv, scale, n_feats, act=False),
            common.default_conv(n_feats, out_feats, kernel_size)]
        self.head = nn.Sequential(*modules_head)
        self.body = nn.Sequential(*modules_body)
        self.tail = nn.Sequential(*modules_tail)
        self._cfg = dict(n_feats=n_feats, n_groups=n_resgroups, n_blocks=n_resblocks, reduction=reduction,
                         scale=scale, res_scale=1.0, in_feats=in_feats, out_feats=out_feats)

    def _engine_kwargs(self):
        return _engine.ARCH_RCAN, dict(self._cfg)


class EDSR(_NativeTrunk):
    """reference architectures.py:198-257"""

    def __init__(self, in_features=3, out_features=3, net_features=64, num_blocks=16, scale=4, res_scale=0.1):
        super(EDSR, self).__init__()
        n_feats = net_features
        kernel_size = 3
        act = nn.ReLU(True)
        m_head = [common.default_conv(in_features, n_feats, kernel_size)]
        m_body = [common.ResBlock(common.default_conv, n_feats, kernel_size, act=act, res_scale=res_scale)
                  for _ in range(num_blocks)]
        m_body.append(common.default_conv(n_feats, n_feats, kernel_size))
        m_tail = [common.Upsampler(common.default_conv, scale, n_feats),
                  common.default_conv(n_feats, out_features, kernel_size)]
        self.head = nn.Sequential(*m_head)
        self.body = nn.Sequential(*m_body)
        self.tail = nn.Sequential(*m_tail)
        self._cfg = dict(n_feats=n_feats, n_groups=1, n_blocks=num_blocks, reduction=16, scale=scale,
                         res_scale=res_scale, in_feats=in_features, out_feats=out_features)

    def _engine_kwargs(self):
        return _engine.ARCH_EDSR, dict(self._cfg)


class HAN(_NativeTrunk):
    """reference architectures.py:331-394: RCAN's residual groups, then layer attention (LAM) over the 11 stacked
    group / body outputs -> last_conv, channel-spatial attention (CSAM) of the body output, last(cat) + head skip.
    The groups run in the trunk kernels; LAM / CSAM (forward and backward) are the kernels of csrc/han.cu."""

    def __init__(self, n_resgroups=10, n_resblocks=20, n_feats=64, reduction=16, scale=4, n_colors=3, res_scale=1.0,
                 conv=common.default_conv):
        super(HAN, self).__init__()
        kernel_size = 3
        act = nn.ReLU(True)
        modules_head = [conv(n_colors, n_feats, kernel_size)]
        modules_body = [
            ResidualGroup(conv, n_feats, kernel_size, reduction, act=act, res_scale=res_scale, n_resblocks=n_resblocks)
            for _ in range(n_resgroups)]
        modules_body.append(conv(n_feats, n_feats, kernel_size))
        modules_tail = [
            common.Upsampler(conv, scale, n_feats, act=False),
            conv(n_feats, n_colors, kernel_size)]
        self.head = nn.Sequential(*modules_head)
        self.body = nn.Sequential(*modules_body)
        self.csa = CSAM_Module(n_feats)
        self.la = LAM_Module(n_feats)
        self.last_conv = nn.Conv2d(n_feats * 11, n_feats, 3, 1, 1)     # the reference fixes 11 = 10 groups + body conv
        self.last = nn.Conv2d(n_feats * 2, n_feats, 3, 1, 1)
        self.tail = nn.Sequential(*modules_tail)
        self._cfg = dict(n_feats=n_feats, n_groups=n_resgroups, n_blocks=n_resblocks, reduction=reduction,
                         scale=scale, res_scale=1.0, in_feats=n_colors, out_feats=n_colors)

    def _engine_kwargs(self):
        return _engine.ARCH_HAN, dict(self._cfg)
