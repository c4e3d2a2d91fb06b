"""B200-native mirror of the reference's Q-RCAN and Q-EDSR (meta-attention RCAN / EDSR)
(/root/reference/rumpy/SISR/models/attention_manipulators/architectures.py:46-556).

Same class names, constructor signatures, module tree, registration order and state_dict keys as the reference
for the configurations the native trunk implements:
  * QCALayer style 'standard' (plain channel attention, :127-128) or 'modulate' (attention x attributes, :113-114);
  * optional q-nodes (`include_q_layer`, `selective_meta_blocks`, `num_q_layers_inner_residual`) with the
    reference's 2-layer ParaCALayer (q_layer.py:5-45).
Everything else the reference's QRCAN can be configured with (pixel attention, SFT / DGFMB / DA-conv layers, the
concat styles, staggered encodings, outer metadata reduction) is outside SURVEY.md section 8 and raises
NotImplementedError at construction.  Q-RCAN style 'standard' and Q-EDSR train natively (the q-layer parameters get
their gradients from per-channel sums the backward kernels leave behind), style 'modulate' included.

`QEDSR` (ParamResBlocks: res_scale * conv2(relu(conv1 x)) * q + x) keeps the reference's ctor and key layout too
(`head.weight` without a Sequential index, `final_body` registered before `body`).

`QRCAN.forward(x, metadata)` / `QEDSR.forward(x, metadata)` hand the parameter list and the metadata to the native executor: the metadata
multipliers q[rcab][n][c] are computed by one kernel, and the whole body runs in the same trunk kernel as RCAN
with q multiplied into the channel-attention vector.
"""
import torch
from torch import nn

from rumpy_b200 import _lib
from rumpy_b200 import engine as _engine
from rumpy_b200.SISR.models.advanced import common
from rumpy_b200.SISR.models.advanced.architectures import _NativeTrunk
from rumpy_b200.SISR.models.advanced.HAN_blocks import CSAM_Module, LAM_Module
from rumpy_b200.SISR.models.attention_manipulators.q_layer import ParaCALayer
from rumpy_b200.trunk_function import trunk_apply

_NATIVE_ONLY = 'runs inside the native QRCAN trunk only (no standalone / CPU path)'


class QCALayer(nn.Module):
    """reference architectures.py:46-130"""

    def __init__(self, channel, style, reduction=16, num_metadata=1):
        super(QCALayer, self).__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        if reduction < 16:
            raise RuntimeError('Using an extreme channel attention reduction value')
        if style not in ('standard', 'modulate'):
            raise NotImplementedError(f"rumpy_b200 QCALayer: style '{style}' (native: 'standard', 'modulate')")
        self.conv_du = nn.Sequential(
            nn.Conv2d(channel, channel // reduction, 1, padding=0, bias=True),
            nn.ReLU(inplace=True),
            nn.Conv2d(channel // reduction, channel, 1, padding=0, bias=True),
            nn.Sigmoid()
        )
        self.style = style

    def forward(self, x, attributes):
        raise _lib.RumpyB200Error('QCALayer ' + _NATIVE_ONLY)


class QRCAB(nn.Module):
    """reference architectures.py:152-219.  Registration order: final_body (channel attention), q_node, body."""

    def __init__(self, conv, n_feat, kernel_size, reduction, style='modulate', pa=False, q_layer=False,
                 dgfmb_layer=False, sft_layer=False, da_conv_layer=False, bias=True, bn=False, act=nn.ReLU(True),
                 res_scale=1, num_metadata=1, num_layers_in_q_layer=2, num_layers_in_dgfmb_layer=2,
                 use_dgfmb_reduction=True):
        super(QRCAB, self).__init__()
        if bn or pa or dgfmb_layer or sft_layer or da_conv_layer:
            raise NotImplementedError('rumpy_b200 QRCAB: bn / pixel attention / DGFMB / SFT / DA-conv variants')
        if q_layer and num_layers_in_q_layer != 2:
            raise NotImplementedError('rumpy_b200 QRCAB: q-layers with 2 fully-connected layers only')
        modules_body = []
        for i in range(2):
            modules_body.append(conv(n_feat, n_feat, kernel_size, bias=bias))
            if i == 0:
                modules_body.append(act)
        self.final_body = QCALayer(channel=n_feat, reduction=reduction, style=style, num_metadata=num_metadata)
        self.pa = pa
        self.q_layer = q_layer
        self.dgfmb_layer = dgfmb_layer
        self.da_conv_layer = da_conv_layer
        self.sft_layer = sft_layer
        if q_layer:
            self.q_node = ParaCALayer(network_channels=n_feat, num_metadata=num_metadata, nonlinearity=True,
                                      num_layers=num_layers_in_q_layer)
        self.body = nn.Sequential(*modules_body)
        self.res_scale = res_scale

    def forward(self, x):
        raise _lib.RumpyB200Error('QRCAB ' + _NATIVE_ONLY)


class QResidualGroup(nn.Module):
    """reference architectures.py:247-300"""

    def __init__(self, conv, n_feat, kernel_size, reduction, act, res_scale, n_resblocks, style, num_metadata,
                 pa, q_layer, dgfmb_layer, da_conv_layer,
                 num_q_layers, num_layers_in_q_layer,
                 sft_layer, num_sft_layers,
                 num_dgfmb_layers, num_layers_in_dgfmb_layer, use_dgfmb_reduction,
                 num_da_conv_layers):
        super(QResidualGroup, self).__init__()
        modules_body = []
        for index in range(n_resblocks):
            q_in = q_layer if (num_q_layers is None or index < num_q_layers) else False
            dgfmb_in = dgfmb_layer if (num_dgfmb_layers is None or index < num_dgfmb_layers) else False
            da_conv_in = da_conv_layer if (num_da_conv_layers is None or index < num_da_conv_layers) else False
            sft_in = sft_layer if (num_sft_layers is None or index < num_sft_layers) else False
            modules_body.append(QRCAB(conv, n_feat, kernel_size, reduction, bias=True, bn=False,
                                      act=act, res_scale=res_scale, style=style,
                                      pa=pa, q_layer=q_in, dgfmb_layer=dgfmb_in, da_conv_layer=da_conv_in,
                                      num_metadata=num_metadata,
                                      num_layers_in_q_layer=num_layers_in_q_layer,
                                      sft_layer=sft_in,
                                      num_layers_in_dgfmb_layer=num_layers_in_dgfmb_layer,
                                      use_dgfmb_reduction=use_dgfmb_reduction))
        self.final_body = conv(n_feat, n_feat, kernel_size)
        self.body = nn.Sequential(*modules_body)

    def forward(self, x):
        raise _lib.RumpyB200Error('QResidualGroup ' + _NATIVE_ONLY)


class QRCAN(_NativeTrunk):
    """reference architectures.py:313-446"""

    def __init__(self, n_resblocks=20, n_resgroups=10, n_feats=64, in_feats=3, out_feats=3, scale=4, reduction=16,
                 res_scale=1.0, style='modulate', num_metadata=1, include_pixel_attention=False,
                 selective_meta_blocks=None,
                 include_q_layer=False, num_q_layers_inner_residual=None, num_layers_in_q_layer=2,
                 include_sft_layer=False, num_sft_layers_inner_residual=None,
                 include_dgfmb_layer=False, num_dgfmb_layers_inner_residual=None, num_layers_in_dgfmb_layer=2,
                 use_dgfmb_reduction=True, use_dgfmb_outer_reduction=False,
                 include_da_conv_layer=False, num_da_conv_layers_inner_residual=None, staggered_encoding=False,
                 **kwargs):
        super(QRCAN, self).__init__()
        kernel_size = 3
        act = nn.ReLU(True)
        if style != 'standard' and staggered_encoding:
            raise RuntimeError('QRCAN must be set to standard for staggered encoding to work.')
        if staggered_encoding or use_dgfmb_outer_reduction:
            raise NotImplementedError('rumpy_b200 QRCAN: staggered encodings / outer metadata reduction')
        self.style = style
        self.staggered_encoding = staggered_encoding
        self.metadata_reduction = nn.Sequential(nn.Identity())

        modules_head = [common.default_conv(in_feats, n_feats, kernel_size)]
        modules_body = []
        for index in range(n_resgroups):
            on = selective_meta_blocks is None or bool(selective_meta_blocks[index])
            modules_body.append(
                QResidualGroup(common.default_conv, n_feats, kernel_size, reduction, style=style,
                               num_metadata=num_metadata, pa=include_pixel_attention,
                               q_layer=include_q_layer if on else False,
                               dgfmb_layer=include_dgfmb_layer if on else False,
                               da_conv_layer=include_da_conv_layer if on else False,
                               sft_layer=include_sft_layer if on else False,
                               act=act, res_scale=res_scale, n_resblocks=n_resblocks,
                               num_q_layers=num_q_layers_inner_residual,
                               num_layers_in_q_layer=num_layers_in_q_layer,
                               num_dgfmb_layers=num_dgfmb_layers_inner_residual,
                               num_layers_in_dgfmb_layer=num_layers_in_dgfmb_layer,
                               num_sft_layers=num_sft_layers_inner_residual,
                               use_dgfmb_reduction=use_dgfmb_reduction,
                               num_da_conv_layers=num_da_conv_layers_inner_residual))
        self.final_body = common.default_conv(n_feats, n_feats, kernel_size)
        modules_tail = [
            common.Upsampler(common.default_conv, scale, n_feats, act=False),
            common.default_conv(n_feats, out_feats, kernel_size)]
        self.head = nn.Sequential(*modules_head)
        self.body = nn.Sequential(*modules_body)
        self.tail = nn.Sequential(*modules_tail)

        has_q = [bool(blk.q_layer) for grp in self.body for blk in grp.body]
        q_hidden = 0
        for grp in self.body:
            for blk in grp.body:
                if blk.q_layer:
                    q_hidden = blk.q_node.layer_sizes[1]
        self._cfg = dict(n_feats=n_feats, n_groups=n_resgroups, n_blocks=n_resblocks, reduction=reduction,
                         scale=scale, in_feats=in_feats, out_feats=out_feats, num_metadata=num_metadata,
                         q_hidden=max(q_hidden, 1), block_has_q=has_q, modulate=(style == 'modulate'))

    def _engine_kwargs(self):
        return _engine.ARCH_QRCAN, dict(self._cfg)

    def forward(self, x, metadata):
        eng = self.native_engine()
        eng.set_metadata(metadata, x.shape[0])
        return trunk_apply(self, x)       # autograd boundary: native backward incl. the q-layer parameters

    def forensic(self, x, qpi, *args, **kwargs):
        raise NotImplementedError('rumpy_b200: forensic() diagnostics are not part of the native trunk')


class ParamResBlock(nn.Module):
    """reference architectures.py:463-493.  Registration order: body, attention_layer."""

    def __init__(self, conv, n_feats, n_params, kernel_size, act=nn.ReLU(True), bias=True, res_scale=1.0,
                 q_layer_nonlinearity=False, add_q_layer=None, num_layers=2):
        super(ParamResBlock, self).__init__()
        if add_q_layer and num_layers != 2:
            raise NotImplementedError('rumpy_b200 ParamResBlock: q-layers with 2 fully-connected layers only')
        m = []
        for i in range(2):
            m.append(conv(n_feats, n_feats, kernel_size, bias=bias))
            if i == 0:
                m.append(act)
        self.body = nn.Sequential(*m)
        self.add_q_layer = add_q_layer
        if self.add_q_layer:
            self.attention_layer = ParaCALayer(n_feats, n_params, nonlinearity=q_layer_nonlinearity,
                                               num_layers=num_layers)
        self.res_scale = res_scale

    def forward(self, x):
        raise _lib.RumpyB200Error('ParamResBlock ' + _NATIVE_ONLY)


class QEDSR(_NativeTrunk):
    """reference architectures.py:496-556"""

    def __init__(self, in_features=3, out_features=3, num_features=64, input_para=1, num_blocks=16, scale=4,
                 res_scale=0.1, q_layer_nonlinearity=False, selective_meta_blocks=None, num_layers=2, **kwargs):
        super(QEDSR, self).__init__()
        n_feats = num_features
        kernel_size = 3
        if selective_meta_blocks == 'front_only':
            selective_meta_blocks = [True] + [False] * (num_blocks - 1)
        self.head = common.default_conv(in_features, n_feats, kernel_size)
        m_body = [
            ParamResBlock(common.default_conv, n_feats, input_para, kernel_size, res_scale=res_scale,
                          q_layer_nonlinearity=q_layer_nonlinearity,
                          add_q_layer=True if selective_meta_blocks is None else selective_meta_blocks[i],
                          num_layers=num_layers) for i in range(num_blocks)]
        self.final_body = common.default_conv(n_feats, n_feats, kernel_size)
        m_tail = [common.Upsampler(common.default_conv, scale, n_feats),
                  common.default_conv(n_feats, out_features, kernel_size)]
        self.body = nn.Sequential(*m_body)
        self.tail = nn.Sequential(*m_tail)
        has_q = [bool(blk.add_q_layer) for blk in self.body]
        q_hidden = max([blk.attention_layer.layer_sizes[1] for blk in self.body if blk.add_q_layer] + [1])
        self._cfg = dict(n_feats=n_feats, n_groups=1, n_blocks=num_blocks, scale=scale, res_scale=res_scale,
                         in_feats=in_features, out_feats=out_features, num_metadata=input_para, q_hidden=q_hidden,
                         block_has_q=has_q, q_relu=bool(q_layer_nonlinearity))

    def _engine_kwargs(self):
        return _engine.ARCH_QEDSR, dict(self._cfg)

    forward = QRCAN.forward


class QHAN(_NativeTrunk):
    """reference architectures.py:643-760: HAN whose residual groups are QResidualGroups (meta-attention inside the
    RCABs), i.e. the Q-RCAN trunk followed by HAN's layer / channel-spatial attention.  Same supported options as QRCAN."""

    def __init__(self, n_resgroups=10, n_resblocks=20, n_feats=64, reduction=16, num_metadata=0,
                 scale=4, n_colors=3, res_scale=1.0, style='standard', include_pixel_attention=False,
                 selective_meta_blocks=None,
                 include_q_layer=False, num_q_layers_inner_residual=None, num_layers_in_q_layer=2,
                 include_sft_layer=False, num_sft_layers_inner_residual=None,
                 include_dgfmb_layer=False, num_dgfmb_layers_inner_residual=None, num_layers_in_dgfmb_layer=2,
                 use_dgfmb_reduction=True, use_dgfmb_outer_reduction=False,
                 include_da_conv_layer=False, num_da_conv_layers_inner_residual=None, **kwargs):
        super(QHAN, self).__init__()
        kernel_size = 3
        act = nn.ReLU(True)
        self.style = style
        modules_head = [common.default_conv(n_colors, n_feats, kernel_size)]
        modules_body = []
        for index in range(n_resgroups):
            on = selective_meta_blocks is None or bool(selective_meta_blocks[index])
            modules_body.append(
                QResidualGroup(common.default_conv, n_feats, kernel_size, reduction, style=style,
                               num_metadata=num_metadata, pa=include_pixel_attention,
                               q_layer=include_q_layer if on else False,
                               # the reference passes dgfmb_layer=None when no selection is given (:676)
                               dgfmb_layer=(include_dgfmb_layer if on else False) if selective_meta_blocks is not None else None,
                               da_conv_layer=include_da_conv_layer if on else False,
                               sft_layer=include_sft_layer if on else False,
                               act=act, res_scale=res_scale, n_resblocks=n_resblocks,
                               num_q_layers=num_q_layers_inner_residual,
                               num_layers_in_q_layer=num_layers_in_q_layer,
                               num_dgfmb_layers=num_dgfmb_layers_inner_residual,
                               num_layers_in_dgfmb_layer=num_layers_in_dgfmb_layer,
                               num_sft_layers=num_sft_layers_inner_residual,
                               use_dgfmb_reduction=use_dgfmb_reduction,
                               num_da_conv_layers=num_da_conv_layers_inner_residual))
        modules_body.append(common.default_conv(n_feats, n_feats, kernel_size))
        modules_tail = [
            common.Upsampler(common.default_conv, scale, n_feats, act=False),
            common.default_conv(n_feats, n_colors, kernel_size)]
        self.head = nn.Sequential(*modules_head)
        self.body = nn.Sequential(*modules_body)
        self.csa = CSAM_Module(n_feats)
        self.la = LAM_Module(n_feats)
        self.last_conv = nn.Conv2d(n_feats * 11, n_feats, 3, 1, 1)
        self.last = nn.Conv2d(n_feats * 2, n_feats, 3, 1, 1)
        self.tail = nn.Sequential(*modules_tail)
        groups = [m for m in self.body if isinstance(m, QResidualGroup)]
        has_q = [bool(blk.q_layer) for grp in groups for blk in grp.body]
        q_hidden = max([blk.q_node.layer_sizes[1] for grp in groups for blk in grp.body if blk.q_layer] + [1])
        self._cfg = dict(n_feats=n_feats, n_groups=n_resgroups, n_blocks=n_resblocks, reduction=reduction,
                         scale=scale, in_feats=n_colors, out_feats=n_colors, num_metadata=max(num_metadata, 1),
                         q_hidden=q_hidden, block_has_q=has_q, modulate=(style == 'modulate'))

    def _engine_kwargs(self):
        return _engine.ARCH_QHAN, dict(self._cfg)

    forward = QRCAN.forward
