"""CPU oracle #2 (TEST INFRASTRUCTURE ONLY) -- functional torch-CPU restatement, fp32.

The reference's arithmetic for this path lives in PyTorch/ATen (requirements.txt:1, unpinned
`pytorch>=1.10.0`; this image: torch 2.11.0+cu128, CPU path = oneDNN).  This module restates the
EDSR/RCAN forward with `torch.nn.functional` calls on the *same* ATen CPU kernels the reference's
`nn.Conv2d` / `nn.AdaptiveAvgPool2d` / `nn.PixelShuffle` modules dispatch to, driven by a plain
state_dict (reference key layout), and gets backward from autograd exactly like the reference's
`loss.backward()` (base_architecture.py:432).  It is therefore the faithful *performance* port of
the reference's CPU path and is what `bench.py` times as `cpu_baseline` / `--impl reference`
(kind "port"), while `oracle/sr_numpy.py` is the independent arithmetic restatement.

Only tests/, __graft_entry__.smoke() and bench.py may import this file; the product never does.

Reference lines restated (relative to /root/reference):
  RCAN.forward  rumpy/SISR/models/advanced/architectures.py:171-176
  ResidualGroup :121-124   RCAB :81-84   CALayer :41-44
  EDSR.forward  :236-241   ResBlock common.py:71-75   Upsampler common.py:29-44
  QRCAN.forward rumpy/SISR/models/attention_manipulators/architectures.py:438-446 (qrcan_forward below)
  HAN.forward   rumpy/SISR/models/advanced/architectures.py:368-392, HAN_blocks.py:17-76 (han_forward below)
  train step    rumpy/shared_framework/models/base_architecture.py:425-440,457-485
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _conv(sd, key, x):
    w = sd[key + '.weight']
    return F.conv2d(x, w, sd[key + '.bias'], padding=w.shape[-1] // 2)


def _ca(sd, key, x):
    y = F.adaptive_avg_pool2d(x, 1)
    y = F.relu(_conv(sd, key + '.conv_du.0', y))
    y = torch.sigmoid(_conv(sd, key + '.conv_du.2', y))
    return x * y


def _tail(sd, x, scale):
    if scale & (scale - 1) == 0:
        for i in range(int(math.log2(scale))):
            x = F.pixel_shuffle(_conv(sd, f'tail.0.{2 * i}', x), 2)
    elif scale == 3:
        x = F.pixel_shuffle(_conv(sd, 'tail.0.0', x), 3)
    else:
        raise NotImplementedError
    return _conv(sd, 'tail.1', x)


def rcan_forward(sd, x, n_resgroups=10, n_resblocks=20, scale=4):
    x = _conv(sd, 'head.0', x)
    res = x
    for g in range(n_resgroups):
        gin = res
        for b in range(n_resblocks):
            p = f'body.{g}.body.{b}.body'
            t = F.relu(_conv(sd, p + '.0', res))
            res = _ca(sd, p + '.3', _conv(sd, p + '.2', t)) + res
        res = _conv(sd, f'body.{g}.body.{n_resblocks}', res) + gin
    res = _conv(sd, f'body.{n_resgroups}', res) + x
    return _tail(sd, res, scale)


def edsr_forward(sd, x, num_blocks=16, res_scale=0.1, scale=4):
    x = _conv(sd, 'head.0', x)
    res = x
    for b in range(num_blocks):
        t = F.relu(_conv(sd, f'body.{b}.body.0', res))
        res = _conv(sd, f'body.{b}.body.2', t).mul(res_scale) + res
    res = _conv(sd, f'body.{num_blocks}', res) + x
    return _tail(sd, res, scale)


def qrcan_forward(sd, x, attributes, n_resgroups, n_resblocks, scale=4, style='standard'):
    """Q-RCAN forward (reference SISR/models/attention_manipulators/architectures.py: QRCAN.forward :438-446,
    QResidualGroup.forward :296-300, QRCAB.forward :198-219 with QCALayer.forward :110-130 styles 'standard' /
    'modulate' and ParaCALayer.forward q_layer.py:42-45).  attributes: [N, M, 1, 1]; RCABs whose state_dict
    holds `q_node.attribute_integrator.*` are multiplied by the meta-attention vector."""
    x = _conv(sd, 'head.0', x)
    res = x
    for g in range(n_resgroups):
        gin = res
        for b in range(n_resblocks):
            p = f'body.{g}.body.{b}'
            u = _conv(sd, p + '.body.2', F.relu(_conv(sd, p + '.body.0', res)))
            y = F.adaptive_avg_pool2d(u, 1)
            y = F.relu(_conv(sd, p + '.final_body.conv_du.0', y))
            y = torch.sigmoid(_conv(sd, p + '.final_body.conv_du.2', y))
            if style == 'modulate':
                y = y * attributes
            u = u * y
            if p + '.q_node.attribute_integrator.0.weight' in sd:
                q = F.relu(_conv(sd, p + '.q_node.attribute_integrator.0', attributes))
                q = torch.sigmoid(_conv(sd, p + '.q_node.attribute_integrator.2', q))
                u = u * q
            res = u + res
        res = _conv(sd, f'body.{g}.final_body', res) + gin
    res = _conv(sd, 'final_body', res) + x
    return _tail(sd, res, scale)


def qedsr_forward(sd, x, attributes, num_blocks, res_scale=0.1, scale=4):
    """Q-EDSR forward (reference attention_manipulators/architectures.py: QEDSR.forward :549-555,
    ParamResBlock.forward :484-493).  The integrator's second conv is `.2` when a ReLU separates the two
    fully-connected layers (q_layer_nonlinearity) and `.1` when it does not."""
    def conv(key, v):
        w = sd[key + '.weight']
        return F.conv2d(v, w, sd[key + '.bias'], padding=w.shape[-1] // 2)
    x = conv('head', x)
    res = x
    for b in range(num_blocks):
        r = conv(f'body.{b}.body.2', F.relu(conv(f'body.{b}.body.0', res))).mul(res_scale)
        p = f'body.{b}.attention_layer.attribute_integrator'
        if p + '.0.weight' in sd:
            q = conv(p + '.0', attributes)
            q = conv(p + '.2', F.relu(q)) if p + '.2.weight' in sd else conv(p + '.1', q)
            r = r * torch.sigmoid(q)
        res = r + res
    res = conv('final_body', res) + x
    return _tail(sd, res, scale)


def qhan_forward(sd, x, attributes, n_resblocks, scale=4, style='standard'):
    """Q-HAN forward (reference attention_manipulators/architectures.py:733-760): HAN's attention tail on a trunk of
    QResidualGroups (qrcan_forward's block)."""
    def block(p, res):
        u = _conv(sd, p + '.body.2', F.relu(_conv(sd, p + '.body.0', res)))
        yv = F.adaptive_avg_pool2d(u, 1)
        yv = torch.sigmoid(_conv(sd, p + '.final_body.conv_du.2', F.relu(_conv(sd, p + '.final_body.conv_du.0', yv))))
        if style == 'modulate':
            yv = yv * attributes
        u = u * yv
        if p + '.q_node.attribute_integrator.0.weight' in sd:
            q = F.relu(_conv(sd, p + '.q_node.attribute_integrator.0', attributes))
            u = u * torch.sigmoid(_conv(sd, p + '.q_node.attribute_integrator.2', q))
        return u + res
    x = _conv(sd, 'head.0', x)
    res = x
    stack = []
    for g in range(10):
        gin = res
        for b in range(n_resblocks):
            res = block(f'body.{g}.body.{b}', res)
        res = _conv(sd, f'body.{g}.final_body', res) + gin
        stack.insert(0, res)
    res = _conv(sd, 'body.10', res)
    stack.insert(0, res)
    return _han_tail(sd, x, res, stack, scale)


def _han_tail(sd, x, out1, stack, scale):
    s = torch.stack(stack, 1)
    B, L, C, H, W = s.shape
    q = s.reshape(B, L, -1)
    energy = torch.bmm(q, q.permute(0, 2, 1))
    att = torch.softmax(energy.max(-1, keepdim=True)[0].expand_as(energy) - energy, dim=-1)
    la = (sd['la.gamma'] * torch.bmm(att, q).reshape(B, L, C, H, W) + s).reshape(B, -1, H, W)
    out2 = _conv(sd, 'last_conv', la)
    a = torch.sigmoid(F.conv3d(out1.unsqueeze(1), sd['csa.conv.weight'], sd['csa.conv.bias'], padding=1))
    a = (sd['csa.gamma'] * a).reshape(B, -1, H, W)
    out1 = out1 * a + out1
    res = _conv(sd, 'last', torch.cat([out1, out2], 1)) + x
    return _tail(sd, res, scale)


def han_forward(sd, x, n_resgroups=10, n_resblocks=20, scale=4):
    """HAN forward (reference advanced/architectures.py:368-392, HAN_blocks.py: LAM_Module.forward :17-41,
    CSAM_Module.forward :55-76)."""
    x = _conv(sd, 'head.0', x)
    res = x
    stack = []
    for g in range(n_resgroups):
        gin = res
        for b in range(n_resblocks):
            p = f'body.{g}.body.{b}.body'
            t = F.relu(_conv(sd, p + '.0', res))
            res = _ca(sd, p + '.3', _conv(sd, p + '.2', t)) + res
        res = _conv(sd, f'body.{g}.body.{n_resblocks}', res) + gin
        stack.insert(0, res)
    res = _conv(sd, f'body.{n_resgroups}', res)
    stack.insert(0, res)
    out1 = res
    s = torch.stack(stack, 1)                                     # B x L x C x H x W, newest first
    B, L, C, H, W = s.shape
    q = s.reshape(B, L, -1)
    energy = torch.bmm(q, q.permute(0, 2, 1))
    att = torch.softmax(energy.max(-1, keepdim=True)[0].expand_as(energy) - energy, dim=-1)
    la = (sd['la.gamma'] * torch.bmm(att, q).reshape(B, L, C, H, W) + s).reshape(B, -1, H, W)
    out2 = _conv(sd, 'last_conv', la)
    a = torch.sigmoid(F.conv3d(out1.unsqueeze(1), sd['csa.conv.weight'], sd['csa.conv.bias'], padding=1))
    a = (sd['csa.gamma'] * a).reshape(B, -1, H, W)
    out1 = out1 * a + out1
    res = _conv(sd, 'last', torch.cat([out1, out2], 1)) + x
    return _tail(sd, res, scale)


def infer_arch(sd):
    """Recover (arch, kwargs) from a reference-layout state_dict."""
    keys = list(sd.keys())
    n_feats = sd['head.0.weight'].shape[0]
    scale_convs = sorted({int(k.split('.')[2]) for k in keys if k.startswith('tail.0.')})
    cout = sd['tail.0.0.weight'].shape[0]
    scale = 3 if cout == 9 * n_feats else 2 ** len(scale_convs)
    if any('.conv_du.' in k for k in keys):
        groups = sorted({int(k.split('.')[1]) for k in keys if k.startswith('body.')})
        n_resgroups = max(groups)
        blocks = {int(k.split('.')[3]) for k in keys if k.startswith('body.0.body.')}
        return 'rcan', dict(n_resgroups=n_resgroups, n_resblocks=max(blocks), n_feats=n_feats, scale=scale)
    blocks = sorted({int(k.split('.')[1]) for k in keys if k.startswith('body.')})
    return 'edsr', dict(num_blocks=max(blocks), n_feats=n_feats, scale=scale)


def forward(sd, x, arch, **kw):
    if arch == 'han':
        return han_forward(sd, x, kw.get('n_resgroups', 10), kw['n_resblocks'], kw.get('scale', 4))
    if arch == 'qedsr':
        return qedsr_forward(sd, x, kw['attributes'], kw['num_blocks'], kw.get('res_scale', 0.1), kw.get('scale', 4))
    if arch == 'qrcan':
        return qrcan_forward(sd, x, kw['attributes'], kw['n_resgroups'], kw['n_resblocks'], kw.get('scale', 4),
                             kw.get('style', 'standard'))
    if arch == 'rcan':
        return rcan_forward(sd, x, kw['n_resgroups'], kw['n_resblocks'], kw.get('scale', 4))
    return edsr_forward(sd, x, kw['num_blocks'], kw.get('res_scale', 0.1), kw.get('scale', 4))


class Trainer:
    """Restates BaseModel.run_train (base_architecture.py:457-485): L1 loss, Adam(lr), optional
    clip_grad_norm_, per-batch scheduler omitted (constant lr unless `lr_fn` given)."""

    def __init__(self, sd, arch, lr=1e-4, grad_clip=None, **kw):
        self.params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        self.arch, self.kw, self.grad_clip = arch, kw, grad_clip
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr)

    def step(self, x, y):
        out = forward(self.params, x, self.arch, **self.kw)
        loss = F.l1_loss(out, y)
        self.opt.zero_grad()
        loss.backward()
        if self.grad_clip is not None:
            torch.nn.utils.clip_grad_norm_(list(self.params.values()), self.grad_clip)
        self.opt.step()
        return float(loss.detach()), out.detach()

    def grads(self):
        return {k: v.grad.detach().clone() for k, v in self.params.items()}
