"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- numpy restatement of RUMpy's EDSR/RCAN trunk.

This file is the parity checker for the CUDA path.  It is *never* imported by the product
(`rumpy_b200/`): only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it.

Every function restates, in plain fp32 numpy, what the reference computes through stock
`torch.nn` modules (the reference has no arithmetic of its own; see DESIGN.md "Oracle").
Citations are relative to /root/reference:

  conv3x3 / conv1x1 ....... rumpy/SISR/models/advanced/common.py:6-9   (nn.Conv2d, padding=k//2)
  CALayer ................. rumpy/SISR/models/advanced/architectures.py:24-44
  RCAB .................... architectures.py:60-84   (res_scale stored, never applied: :79)
  ResidualGroup ........... architectures.py:107-124
  RCAN .................... architectures.py:140-176
  ResBlock ................ common.py:51-75         (.mul(res_scale) then += x)
  EDSR .................... architectures.py:198-241
  Upsampler/PixelShuffle .. common.py:23-48
  L1 loss ................. rumpy/shared_framework/models/base_architecture.py:40 (nn.L1Loss, mean)
  Adam .................... base_architecture.py:93-95 (torch.optim.Adam defaults)
  train step .............. base_architecture.py:425-440, 457-485

Parity pin: the reference's own tests hold no numeric vectors for this path (shape checks only,
automated_testing/sisr_tests/test_model_cpu_execute.py:42-49), so this oracle is pinned against
outputs of the reference itself, generated in the build container by
`tests/golden/make_golden.py` and committed under `tests/golden/*.npz`
(see tests/test_oracle_golden.py).

Tensors are NCHW float32, weights OIHW float32, state_dict keys are the reference's.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# primitive ops (forward)
# --------------------------------------------------------------------------------------
def _im2col3(x: np.ndarray) -> np.ndarray:
    """x [N,C,H,W] -> cols [N, C*9, H*W] with zero padding 1 (index = c*9 + ky*3 + kx)."""
    n, c, h, w = x.shape
    xp = np.zeros((n, c, h + 2, w + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    cols = np.empty((n, c, 9, h, w), dtype=x.dtype)
    for ky in range(3):
        for kx in range(3):
            cols[:, :, ky * 3 + kx] = xp[:, :, ky:ky + h, kx:kx + w]
    return cols.reshape(n, c * 9, h * w)


def conv2d(x: np.ndarray, weight: np.ndarray, bias: np.ndarray | None) -> np.ndarray:
    """Cross-correlation, stride 1, zero pad k//2 (common.py:6-9).  k in {1,3}."""
    n, c, h, w = x.shape
    o, ci, kh, kw = weight.shape
    assert ci == c and kh == kw and kh in (1, 3)
    if kh == 1:
        out = np.matmul(weight.reshape(o, c), x.reshape(n, c, h * w))
    else:
        out = np.matmul(weight.reshape(o, c * 9), _im2col3(x))
    out = out.reshape(n, o, h, w)
    if bias is not None:
        out = out + bias.reshape(1, o, 1, 1)
    return out.astype(F32, copy=False)


def conv2d_backward(x, weight, g, need_dx=True):
    """Returns (dx, dw, db) for conv2d above.  SURVEY 8(a') row 1."""
    n, c, h, w = x.shape
    o = weight.shape[0]
    k = weight.shape[2]
    g2 = g.reshape(n, o, h * w)
    db = g2.sum(axis=(0, 2)).astype(F32)
    if k == 1:
        dw = np.einsum('nop,ncp->oc', g2, x.reshape(n, c, h * w), optimize=True).reshape(o, c, 1, 1)
        dx = np.matmul(weight.reshape(o, c).T, g2).reshape(n, c, h, w) if need_dx else None
        return dx, dw.astype(F32), db
    cols = _im2col3(x)
    dw = np.einsum('nop,nkp->ok', g2, cols, optimize=True).reshape(o, c, 3, 3).astype(F32)
    dx = None
    if need_dx:
        # dX = conv of g with 180-degree-rotated, in/out swapped weights
        wt = np.ascontiguousarray(weight[:, :, ::-1, ::-1].transpose(1, 0, 2, 3))
        dx = conv2d(g, wt, None)
    return dx, dw, db


def pixel_shuffle(x: np.ndarray, r: int) -> np.ndarray:
    """out[n,c,h*r+i,w*r+j] = in[n,c*r*r+i*r+j,h,w]  (nn.PixelShuffle, common.py:33,40)."""
    n, crr, h, w = x.shape
    c = crr // (r * r)
    return x.reshape(n, c, r, r, h, w).transpose(0, 1, 4, 2, 5, 3).reshape(n, c, h * r, w * r)


def pixel_unshuffle(g: np.ndarray, r: int) -> np.ndarray:
    n, c, hr, wr = g.shape
    h, w = hr // r, wr // r
    return g.reshape(n, c, h, r, w, r).transpose(0, 1, 3, 5, 2, 4).reshape(n, c * r * r, h, w)


def sigmoid(z):
    return (1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(F32)


# --------------------------------------------------------------------------------------
# network description helpers
# --------------------------------------------------------------------------------------
def upsampler_stages(scale: int):
    """Upsampler structure (common.py:29-44): list of (sequential index of conv, shuffle r)."""
    if scale & (scale - 1) == 0:
        return [(2 * i, 2) for i in range(int(math.log2(scale)))]
    if scale == 3:
        return [(0, 3)]
    raise NotImplementedError(scale)


class _Tape:
    """Minimal tape: forward pushes closures, backward pops them in reverse."""

    def __init__(self):
        self.ops = []


# --------------------------------------------------------------------------------------
# forward (+ optional saved state for backward)
# --------------------------------------------------------------------------------------
def _conv(sd, key, x, saved=None):
    if saved is not None:
        saved.append(('conv', key, x))
    return conv2d(x, sd[key + '.weight'], sd[key + '.bias'])


def ca_layer(sd, key, x, saved=None):
    """architectures.py:41-44."""
    n, c, h, w = x.shape
    mean = x.mean(axis=(2, 3), dtype=np.float64).astype(F32)                  # [N,C]
    w1 = sd[key + '.conv_du.0.weight'].reshape(-1, c)
    b1 = sd[key + '.conv_du.0.bias']
    w2 = sd[key + '.conv_du.2.weight'].reshape(c, -1)
    b2 = sd[key + '.conv_du.2.bias']
    z1 = mean @ w1.T + b1
    hid = np.maximum(z1, 0).astype(F32)
    y = sigmoid(hid @ w2.T + b2)                                              # [N,C]
    if saved is not None:
        saved.append(('ca', key, x, mean, hid, y))
    return (x * y[:, :, None, None]).astype(F32)


def rcan_forward(sd, x, n_resgroups, n_resblocks, scale=4, saved=None):
    """architectures.py:171-176."""
    x = _conv(sd, 'head.0', x, saved)
    head = x
    res = x
    for g in range(n_resgroups):
        gin = res
        for b in range(n_resblocks):
            bin_ = res
            p = f'body.{g}.body.{b}.body'
            t = _conv(sd, p + '.0', res, saved)
            if saved is not None:
                saved.append(('relu', t > 0))
            t = np.maximum(t, 0)
            u = _conv(sd, p + '.2', t, saved)
            res = ca_layer(sd, p + '.3', u, saved) + bin_                     # RCAB: res += x (:83)
        res = _conv(sd, f'body.{g}.body.{n_resblocks}', res, saved) + gin     # group skip (:123)
    res = _conv(sd, f'body.{n_resgroups}', res, saved) + head                 # global skip (:174)
    return _tail(sd, res, scale, saved)


def edsr_forward(sd, x, num_blocks, res_scale, scale=4, saved=None):
    """architectures.py:236-241 with ResBlock common.py:71-75."""
    x = _conv(sd, 'head.0', x, saved)
    head = x
    res = x
    for b in range(num_blocks):
        bin_ = res
        t = _conv(sd, f'body.{b}.body.0', res, saved)
        if saved is not None:
            saved.append(('relu', t > 0))
        t = np.maximum(t, 0)
        u = _conv(sd, f'body.{b}.body.2', t, saved)
        if saved is not None:
            saved.append(('scale', F32(res_scale)))
        res = (u * F32(res_scale) + bin_).astype(F32)
    res = _conv(sd, f'body.{num_blocks}', res, saved) + head
    return _tail(sd, res, scale, saved)


def _tail(sd, x, scale, saved):
    for idx, r in upsampler_stages(scale):
        x = _conv(sd, f'tail.0.{idx}', x, saved)
        if saved is not None:
            saved.append(('shuffle', r))
        x = pixel_shuffle(x, r)
    return _conv(sd, 'tail.1', x, saved)


# --------------------------------------------------------------------------------------
# loss / backward / optimiser
# --------------------------------------------------------------------------------------
def l1_loss(out, y):
    """nn.L1Loss() mean; returns (loss, dOut)."""
    d = out.astype(np.float64) - y.astype(np.float64)
    loss = F32(np.abs(d).mean())
    g = (np.sign(d) / d.size).astype(F32)
    return loss, g


def _ca_backward(sd, key, x, mean, hid, y, g, grads):
    """SURVEY 8(a') CALayer row."""
    n, c, h, w = x.shape
    w1 = sd[key + '.conv_du.0.weight'].reshape(-1, c)
    w2 = sd[key + '.conv_du.2.weight'].reshape(c, -1)
    s = (g.astype(np.float64) * x).sum(axis=(2, 3)).astype(F32)               # dy [N,C]
    dz2 = s * y * (1 - y)
    grads[key + '.conv_du.2.weight'] = (dz2.T @ hid).reshape(c, -1, 1, 1).astype(F32)
    grads[key + '.conv_du.2.bias'] = dz2.sum(0).astype(F32)
    dh = (dz2 @ w2) * (hid > 0)
    grads[key + '.conv_du.0.weight'] = (dh.T @ mean).reshape(-1, c, 1, 1).astype(F32)
    grads[key + '.conv_du.0.bias'] = dh.sum(0).astype(F32)
    dmean = dh @ w1                                                           # [N,C]
    return (g * y[:, :, None, None] + (dmean / F32(h * w))[:, :, None, None]).astype(F32)


def backward(sd, saved, g, arch, n_outer, n_inner=0):
    """Walks the saved tape in reverse.  arch in {'rcan','edsr'}.  Returns grads dict."""
    grads = OrderedDict()
    tape = list(saved)

    def conv_bwd(g, need_dx=True):
        kind, key, x = tape.pop()
        assert kind == 'conv', kind
        dx, dw, db = conv2d_backward(x, sd[key + '.weight'], g, need_dx)
        grads[key + '.weight'] = dw
        grads[key + '.bias'] = db
        return dx

    # tail
    g = conv_bwd(g)
    while tape and tape[-1][0] == 'shuffle':
        _, r = tape.pop()
        g = pixel_unshuffle(g, r)
        g = conv_bwd(g)
    # body tail conv + global skip
    g_head_skip = g
    g = conv_bwd(g)
    if arch == 'rcan':
        for _g in range(n_outer):
            g_group_skip = g
            g = conv_bwd(g)
            for _b in range(n_inner):
                g_blk_skip = g
                kind, key, x, mean, hid, y = tape.pop()
                assert kind == 'ca'
                g = _ca_backward(sd, key, x, mean, hid, y, g, grads)
                g = conv_bwd(g)
                _, mask = tape.pop()
                g = g * mask
                g = conv_bwd(g) + g_blk_skip
            g = g + g_group_skip
    else:
        for _b in range(n_outer):
            g_blk_skip = g
            _, rs = tape.pop()
            g = g * rs
            g = conv_bwd(g)
            _, mask = tape.pop()
            g = g * mask
            g = conv_bwd(g) + g_blk_skip
    g = g + g_head_skip
    conv_bwd(g, need_dx=False)
    assert not tape
    return grads


def adam_step(params, grads, state, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (base_architecture.py:93-95), no weight decay, no amsgrad."""
    state['step'] = state.get('step', 0) + 1
    t = state['step']
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    for k, p in params.items():
        gk = grads[k]
        m = state.setdefault('m.' + k, np.zeros_like(p))
        v = state.setdefault('v.' + k, np.zeros_like(p))
        m *= F32(beta1)
        m += F32(1 - beta1) * gk
        v *= F32(beta2)
        v += F32(1 - beta2) * gk * gk
        denom = np.sqrt(v) / F32(math.sqrt(bc2)) + F32(eps)
        p -= (F32(lr / bc1) * m / denom).astype(F32)


class Net:
    """Convenience wrapper: Net('rcan', n_resgroups=.., n_resblocks=..) or Net('edsr', num_blocks=.., res_scale=..)."""

    def __init__(self, arch, scale=4, **kw):
        self.arch, self.scale, self.kw = arch, scale, kw

    def forward(self, sd, x, saved=None):
        x = np.ascontiguousarray(x, dtype=F32)
        if self.arch == 'rcan':
            return rcan_forward(sd, x, self.kw['n_resgroups'], self.kw['n_resblocks'], self.scale, saved)
        return edsr_forward(sd, x, self.kw['num_blocks'], self.kw.get('res_scale', 0.1), self.scale, saved)

    def loss_and_grads(self, sd, x, y):
        saved = []
        out = self.forward(sd, x, saved)
        loss, g = l1_loss(out, y)
        if self.arch == 'rcan':
            grads = backward(sd, saved, g, 'rcan', self.kw['n_resgroups'], self.kw['n_resblocks'])
        else:
            grads = backward(sd, saved, g, 'edsr', self.kw['num_blocks'])
        return out, loss, grads

    def train_step(self, sd, x, y, state, lr=1e-4):
        out, loss, grads = self.loss_and_grads(sd, x, y)
        adam_step(sd, grads, state, lr=lr)
        return loss, out


def psnr(a, b, max_value=1.0):
    """rumpy/sr_tools/metrics.py:33-44 (mse==0 -> 100)."""
    mse = np.mean((a.astype(F32) - b.astype(F32)) ** 2, dtype=np.float64)
    if mse == 0:
        return 100.0
    return float(20 * math.log10(max_value / math.sqrt(mse)))


def rgb_to_y(img):
    """Y of jpg-style YCbCr on [0,1] RGB, NCHW (image_functions.py:72-88; 0.299/0.587/0.114, no offset)."""
    return 0.299 * img[:, 0] + 0.587 * img[:, 1] + 0.114 * img[:, 2]
