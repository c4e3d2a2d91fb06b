"""Deterministic weights / inputs shared by the golden generator and the tests.

Everything here is numpy-RandomState driven so the same tensors can be rebuilt on any box
(the GPU box has no /root/reference).  Key names / shapes / order follow the reference's
state_dict layout (rumpy/SISR/models/advanced/architectures.py:140-241).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np


def _conv_spec(spec, key, cout, cin, k):
    spec.append((key + '.weight', (cout, cin, k, k)))
    spec.append((key + '.bias', (cout,)))


def _tail_spec(spec, n_feats, out_feats, scale):
    if scale & (scale - 1) == 0:
        for i in range(int(math.log2(scale))):
            _conv_spec(spec, f'tail.0.{2 * i}', 4 * n_feats, n_feats, 3)
    elif scale == 3:
        _conv_spec(spec, 'tail.0.0', 9 * n_feats, n_feats, 3)
    else:
        raise NotImplementedError(scale)
    _conv_spec(spec, 'tail.1', out_feats, n_feats, 3)


def rcan_spec(n_resgroups=10, n_resblocks=20, n_feats=64, reduction=16, scale=4, in_feats=3, out_feats=3):
    spec = []
    _conv_spec(spec, 'head.0', n_feats, in_feats, 3)
    for g in range(n_resgroups):
        for b in range(n_resblocks):
            p = f'body.{g}.body.{b}.body'
            _conv_spec(spec, p + '.0', n_feats, n_feats, 3)
            _conv_spec(spec, p + '.2', n_feats, n_feats, 3)
            _conv_spec(spec, p + '.3.conv_du.0', n_feats // reduction, n_feats, 1)
            _conv_spec(spec, p + '.3.conv_du.2', n_feats, n_feats // reduction, 1)
        _conv_spec(spec, f'body.{g}.body.{n_resblocks}', n_feats, n_feats, 3)
    _conv_spec(spec, f'body.{n_resgroups}', n_feats, n_feats, 3)
    _tail_spec(spec, n_feats, out_feats, scale)
    return spec


def edsr_spec(num_blocks=16, n_feats=64, scale=4, in_feats=3, out_feats=3):
    spec = []
    _conv_spec(spec, 'head.0', n_feats, in_feats, 3)
    for b in range(num_blocks):
        _conv_spec(spec, f'body.{b}.body.0', n_feats, n_feats, 3)
        _conv_spec(spec, f'body.{b}.body.2', n_feats, n_feats, 3)
    _conv_spec(spec, f'body.{num_blocks}', n_feats, n_feats, 3)
    _tail_spec(spec, n_feats, out_feats, scale)
    return spec


def make_weights(spec, seed, gain=1.0):
    """U(-b, b), b = gain/sqrt(fan_in): the distribution of nn.Conv2d's default init."""
    rs = np.random.RandomState(seed)
    sd = OrderedDict()
    fan_in = None
    for key, shape in spec:
        if key == 'weight' or key.endswith('.weight'):
            fan_in = shape[1] * shape[2] * shape[3]
        b = gain / math.sqrt(fan_in)
        sd[key] = rs.uniform(-b, b, size=shape).astype(np.float32)
    return sd


def subsample(a, limit=4096):
    """Deterministic strided subsample used to keep golden gradient fixtures small."""
    flat = np.asarray(a).reshape(-1)
    return flat[::max(1, flat.size // limit)]


def make_input(shape, seed):
    return np.random.RandomState(seed).uniform(0.0, 1.0, size=shape).astype(np.float32)


# name -> (arch, kwargs, lr-input shape, weight seed, input seed)
CASES = OrderedDict(
    rcan_small=('rcan', dict(n_resgroups=2, n_resblocks=2, n_feats=64, scale=4), (2, 3, 12, 20), 11, 12),
    rcan_x2=('rcan', dict(n_resgroups=1, n_resblocks=1, n_feats=64, scale=2), (1, 3, 7, 9), 21, 22),
    rcan_x3=('rcan', dict(n_resgroups=1, n_resblocks=2, n_feats=64, scale=3), (1, 3, 10, 6), 31, 32),
    edsr_small=('edsr', dict(num_blocks=3, n_feats=64, scale=4, res_scale=0.1), (2, 3, 9, 17), 41, 42),
    edsr_wide=('edsr', dict(num_blocks=1, n_feats=256, scale=2, res_scale=0.1), (1, 3, 8, 16), 51, 52),
)


def case_spec(name):
    arch, kw, shape, wseed, xseed = CASES[name]
    if arch == 'rcan':
        spec = rcan_spec(kw['n_resgroups'], kw['n_resblocks'], kw['n_feats'], 16, kw['scale'])
    else:
        spec = edsr_spec(kw['num_blocks'], kw['n_feats'], kw['scale'])
    return arch, kw, shape, spec, wseed, xseed


def case_tensors(name):
    arch, kw, shape, spec, wseed, xseed = case_spec(name)
    sd = make_weights(spec, wseed)
    x = make_input(shape, xseed)
    s = kw['scale']
    y = make_input((shape[0], 3, shape[2] * s, shape[3] * s), xseed + 1000)
    return arch, kw, sd, x, y


# ---------------------------------------------------------------------------- Q-RCAN (meta-attention)
def q_layer_sizes(num_metadata, n_feats=64):
    """ParaCALayer's two layer widths (reference attention_manipulators/q_layer.py:24-31, num_layers=2)."""
    if num_metadata > 15:
        return (n_feats - num_metadata) // 2 + num_metadata, (n_feats - num_metadata) // 1 + num_metadata
    return n_feats // 2, n_feats


def qrcan_has_q(n_resgroups, n_resblocks, include_q_layer, selective_meta_blocks=None,
                num_q_layers_inner_residual=None):
    """Which RCABs own a q-node (reference attention_manipulators/architectures.py:262-266, 384-425)."""
    flags = []
    for g in range(n_resgroups):
        on = include_q_layer and (selective_meta_blocks is None or bool(selective_meta_blocks[g]))
        for b in range(n_resblocks):
            flags.append(bool(on and (num_q_layers_inner_residual is None or b < num_q_layers_inner_residual)))
    return flags


def qrcan_spec(n_resgroups, n_resblocks, num_metadata, has_q, n_feats=64, reduction=16, scale=4, in_feats=3,
               out_feats=3):
    """state_dict layout of the reference's QRCAN (module REGISTRATION order: final_body before head/body/tail,
    and inside a QRCAB the attention and the q-node before the convolutions)."""
    spec = []
    _conv_spec(spec, 'final_body', n_feats, n_feats, 3)
    _conv_spec(spec, 'head.0', n_feats, in_feats, 3)
    h1, h2 = q_layer_sizes(num_metadata, n_feats)
    for g in range(n_resgroups):
        _conv_spec(spec, f'body.{g}.final_body', n_feats, n_feats, 3)
        for b in range(n_resblocks):
            p = f'body.{g}.body.{b}'
            _conv_spec(spec, p + '.final_body.conv_du.0', n_feats // reduction, n_feats, 1)
            _conv_spec(spec, p + '.final_body.conv_du.2', n_feats, n_feats // reduction, 1)
            if has_q[g * n_resblocks + b]:
                _conv_spec(spec, p + '.q_node.attribute_integrator.0', h1, num_metadata, 1)
                _conv_spec(spec, p + '.q_node.attribute_integrator.2', h2, h1, 1)
            _conv_spec(spec, p + '.body.0', n_feats, n_feats, 3)
            _conv_spec(spec, p + '.body.2', n_feats, n_feats, 3)
    _tail_spec(spec, n_feats, out_feats, scale)
    return spec


# name -> (QRCAN ctor kwargs, lr-input shape, weight seed, input seed).  'modulate' cases feed the [N,1] quality
# index through QRCANHandler.scale_qpi (min_mu=-0.2, max_mu=0.8) to get the [N,64] attributes, as the handler does.
QCASES = OrderedDict(
    qrcan_blur_q=(dict(n_resgroups=2, n_resblocks=2, scale=4, style='standard', num_metadata=10,
                       include_q_layer=True), (2, 3, 12, 20), 61, 62),
    qrcan_selective=(dict(n_resgroups=2, n_resblocks=3, scale=2, style='standard', num_metadata=3,
                          include_q_layer=True, selective_meta_blocks=[True, False],
                          num_q_layers_inner_residual=2), (3, 3, 9, 14), 63, 64),
    qrcan_wide_meta=(dict(n_resgroups=1, n_resblocks=2, scale=4, style='standard', num_metadata=40,
                          include_q_layer=True), (2, 3, 10, 10), 65, 66),
    qrcan_modulate=(dict(n_resgroups=1, n_resblocks=2, scale=3, style='modulate', num_metadata=1),
                    (2, 3, 8, 11), 67, 68),
)


def qcase_tensors(name):
    kw, shape, wseed, xseed = QCASES[name]
    has_q = qrcan_has_q(kw['n_resgroups'], kw['n_resblocks'], kw.get('include_q_layer', False),
                        kw.get('selective_meta_blocks'), kw.get('num_q_layers_inner_residual'))
    spec = qrcan_spec(kw['n_resgroups'], kw['n_resblocks'], kw['num_metadata'], has_q, scale=kw['scale'])
    sd = make_weights(spec, wseed)
    x = make_input(shape, xseed)
    meta = make_input((shape[0], kw['num_metadata']), xseed + 500)
    return kw, has_q, sd, x, meta


def qedsr_spec(num_blocks, num_metadata, has_q, n_feats=64, scale=4, q_relu=False, in_feats=3, out_feats=3):
    """state_dict layout of the reference's QEDSR (head is a bare conv; final_body registered before body; the
    integrator's second conv sits at Sequential index 2 only when a ReLU separates the two)."""
    spec = []
    _conv_spec(spec, 'head', n_feats, in_feats, 3)
    _conv_spec(spec, 'final_body', n_feats, n_feats, 3)
    h1, h2 = q_layer_sizes(num_metadata, n_feats)
    for b in range(num_blocks):
        _conv_spec(spec, f'body.{b}.body.0', n_feats, n_feats, 3)
        _conv_spec(spec, f'body.{b}.body.2', n_feats, n_feats, 3)
        if has_q[b]:
            _conv_spec(spec, f'body.{b}.attention_layer.attribute_integrator.0', h1, num_metadata, 1)
            _conv_spec(spec, f'body.{b}.attention_layer.attribute_integrator.{2 if q_relu else 1}', h2, h1, 1)
    _tail_spec(spec, n_feats, out_feats, scale)
    return spec


# name -> (QEDSR ctor kwargs, lr-input shape, weight seed, input seed)
QECASES = OrderedDict(
    qedsr_blur=(dict(num_blocks=3, num_features=64, scale=4, res_scale=0.1, input_para=10,
                     q_layer_nonlinearity=True), (2, 3, 11, 15), 71, 72),
    qedsr_linear_front=(dict(num_blocks=3, num_features=64, scale=2, res_scale=0.1, input_para=2,
                             q_layer_nonlinearity=False, selective_meta_blocks='front_only'), (3, 3, 8, 8), 73, 74),
    qedsr_wide=(dict(num_blocks=1, num_features=256, scale=2, res_scale=0.1, input_para=10,
                     q_layer_nonlinearity=True), (2, 3, 8, 16), 75, 76),
)


def qecase_tensors(name):
    kw, shape, wseed, xseed = QECASES[name]
    sel = kw.get('selective_meta_blocks')
    nb = kw['num_blocks']
    has_q = [True] * nb if sel is None else ([True] + [False] * (nb - 1) if sel == 'front_only' else list(sel))
    spec = qedsr_spec(nb, kw['input_para'], has_q, kw['num_features'], kw['scale'], kw['q_layer_nonlinearity'])
    sd = make_weights(spec, wseed)
    x = make_input(shape, xseed)
    meta = make_input((shape[0], kw['input_para']), xseed + 500)
    return kw, has_q, sd, x, meta


# ---------------------------------------------------------------------------- HAN
def han_spec(n_resblocks, n_feats=64, reduction=16, scale=4):
    """state_dict layout of the reference's HAN (10 residual groups fixed by last_conv's n_feats*11 input)."""
    spec = rcan_spec(10, n_resblocks, n_feats, reduction, scale)
    cut = [i for i, (k, _) in enumerate(spec) if k.startswith('tail.')][0]
    extra = [('csa.gamma', (1,)), ('csa.conv.weight', (1, 1, 3, 3, 3)), ('csa.conv.bias', (1,)), ('la.gamma', (1,))]
    _conv_spec(extra, 'last_conv', n_feats, n_feats * 11, 3)
    _conv_spec(extra, 'last', n_feats, n_feats * 2, 3)
    return spec[:cut] + extra + spec[cut:]


# name -> (n_resblocks, scale, lr-input shape, weight seed, input seed, la.gamma, csa.gamma)
HCASES = OrderedDict(
    han_b1_x4=(1, 4, (2, 3, 12, 20), 81, 82, 0.4, 0.6),
    han_b2_x2=(2, 2, (1, 3, 9, 13), 83, 84, -0.3, 1.0),
)


def hcase_tensors(name):
    nb, scale, shape, wseed, xseed, la_g, csa_g = HCASES[name]
    spec = han_spec(nb, scale=scale)
    rs_fix = {'csa.gamma': csa_g, 'la.gamma': la_g}
    # conv3d fan-in: give the 27-tap kernel O(1) weights so the sigmoid moves
    sd = make_weights([(k, (s if k not in rs_fix else (1, 1, 1, 1))) for k, s in spec], wseed)
    for k, v in rs_fix.items():
        sd[k] = np.full((1,), v, dtype=np.float32)
    sd['csa.conv.weight'] = np.random.RandomState(wseed + 1).uniform(-0.5, 0.5, (1, 1, 3, 3, 3)).astype(np.float32)
    sd['csa.conv.bias'] = np.array([0.1], dtype=np.float32)
    return nb, scale, sd, make_input(shape, xseed)


# ---------------------------------------------------------------------------- Q-HAN (HAN with Q-RCAN's residual groups)
def qhan_spec(n_resblocks, num_metadata, has_q, n_feats=64, reduction=16, scale=4):
    """state_dict layout of the reference's QHAN: head | per group: final_body, per block: attention, [q_node], convs |
    body conv (last entry of `body`) | csa | la | last_conv | last | tail."""
    spec = []
    _conv_spec(spec, 'head.0', n_feats, 3, 3)
    h1, h2 = q_layer_sizes(num_metadata, n_feats)
    for g in range(10):
        _conv_spec(spec, f'body.{g}.final_body', n_feats, n_feats, 3)
        for b in range(n_resblocks):
            p = f'body.{g}.body.{b}'
            _conv_spec(spec, p + '.final_body.conv_du.0', n_feats // reduction, n_feats, 1)
            _conv_spec(spec, p + '.final_body.conv_du.2', n_feats, n_feats // reduction, 1)
            if has_q[g * n_resblocks + b]:
                _conv_spec(spec, p + '.q_node.attribute_integrator.0', h1, num_metadata, 1)
                _conv_spec(spec, p + '.q_node.attribute_integrator.2', h2, h1, 1)
            _conv_spec(spec, p + '.body.0', n_feats, n_feats, 3)
            _conv_spec(spec, p + '.body.2', n_feats, n_feats, 3)
    _conv_spec(spec, 'body.10', n_feats, n_feats, 3)
    spec += [('csa.gamma', (1,)), ('csa.conv.weight', (1, 1, 3, 3, 3)), ('csa.conv.bias', (1,)), ('la.gamma', (1,))]
    _conv_spec(spec, 'last_conv', n_feats, n_feats * 11, 3)
    _conv_spec(spec, 'last', n_feats, n_feats * 2, 3)
    _tail_spec(spec, n_feats, 3, scale)
    return spec


QHCASE = dict(kw=dict(n_resblocks=1, scale=2, style='standard', num_metadata=10, include_q_layer=True,
                      selective_meta_blocks=[True, False] * 5), shape=(2, 3, 10, 12), wseed=91, xseed=92)


def qhcase_tensors():
    kw = QHCASE['kw']
    has_q = qrcan_has_q(10, kw['n_resblocks'], True, kw['selective_meta_blocks'], None)
    spec = qhan_spec(kw['n_resblocks'], kw['num_metadata'], has_q, scale=kw['scale'])
    fix = {'csa.gamma': 0.6, 'la.gamma': 0.4}
    sd = make_weights([(k, (s if k not in fix else (1, 1, 1, 1))) for k, s in spec], QHCASE['wseed'])
    for k, v in fix.items():
        sd[k] = np.full((1,), v, dtype=np.float32)
    sd['csa.conv.weight'] = np.random.RandomState(QHCASE['wseed'] + 1).uniform(-0.5, 0.5, (1, 1, 3, 3, 3)).astype(np.float32)
    sd['csa.conv.bias'] = np.array([0.1], dtype=np.float32)
    x = make_input(QHCASE['shape'], QHCASE['xseed'])
    meta = make_input((QHCASE['shape'][0], kw['num_metadata']), QHCASE['xseed'] + 500)
    return kw, has_q, sd, x, meta
