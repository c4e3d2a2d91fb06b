"""Stand-alone (inference) forwards of the reference's blocks through the C ABI.

The reference's blocks take and return fp32 NCHW tensors (common.py / architectures.py); these helpers
convert at the block boundary, keep NHWC bf16 + fp32 tensors inside, and call the sm_100a kernels.
Whole RCAN / EDSR networks do NOT come through here (see engine.py) -- this is the path for a block used on
its own.  No autograd at block level: training goes through the whole-network executor.
"""
from __future__ import annotations

import torch

from . import _lib, ops


def _require_cuda(x):
    if not x.is_cuda:
        raise _lib.RumpyB200Error('rumpy_b200 has no CPU path: pass CUDA tensors (sm_100 device)')
    return x.contiguous().float()


def _packed(conv, shuffle_r=1, rows_padded=0):
    """Packs (and caches on the module) the conv's bf16 operand; refreshed when the weight changes."""
    from .engine import PARAM_EPOCH
    key = (conv.weight._version, PARAM_EPOCH[0], conv.weight.data_ptr(), shuffle_r, rows_padded)
    if getattr(conv, '_rb_pack_key', None) != key:
        w = conv.weight.detach()
        conv._rb_w = ops.pack_conv3x3(w, rows_padded=rows_padded, shuffle_r=shuffle_r)
        b = conv.bias.detach() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
        conv._rb_b = ops.pack_bias(b, rows_padded=rows_padded, shuffle_r=shuffle_r)
        conv._rb_pack_key = key
    return conv._rb_w, conv._rb_b


def conv_nhwc(conv, xb, N, H, W, **kw):
    cout, cin = conv.weight.shape[:2]
    wp, bp = _packed(conv, kw.get('out_shuffle_r', 1))
    ops.conv3x3(xb, wp, bp, N=N, H=H, W=W, Cin=cin, Cout=cout, **kw)


def conv_forward(conv, x):
    """default_conv / nn.Conv2d call site (reference common.py:6-9)."""
    x = _require_cuda(x)
    N, cin, H, W = x.shape
    cout, _, kh, kw_ = conv.weight.shape
    if (kh, kw_) != (3, 3) or conv.stride != (1, 1) or conv.padding != (1, 1):
        raise NotImplementedError('rumpy_b200 Conv2d: only 3x3 / stride 1 / pad 1 convolutions are on the path '
                                  '(1x1 convs live inside CALayer)')
    bias = conv.bias.detach() if conv.bias is not None else torch.zeros(cout, device=x.device)
    if cin <= 4:
        yf = torch.empty((N, H, W, cout), dtype=torch.float32, device=x.device)
        yb = torch.empty((N, H, W, cout), dtype=torch.bfloat16, device=x.device)
        ops.head_conv(x, conv.weight.detach().contiguous(), bias, yf, yb)
        return ops.nhwc_to_nchw(yf)
    _, xb = ops.nchw_to_nhwc(x, want_f32=False)
    if cout <= 16:
        wp, bp = _packed(conv, 1, 16)
        y = torch.empty((N, cout, H, W), dtype=torch.float32, device=x.device)
        ops.conv3x3_tail(xb, wp, bp, y, N=N, H=H, W=W, Cin=cin, cout_real=cout)
        return y
    yf = torch.empty((N, H, W, cout), dtype=torch.float32, device=x.device)
    conv_nhwc(conv, xb, N, H, W, out_f32=yf)
    return ops.nhwc_to_nchw(yf)


def resblock_forward(blk, x):
    """ResBlock.forward (reference common.py:71-75)."""
    x = _require_cuda(x)
    N, C, H, W = x.shape
    xf, xb = ops.nchw_to_nhwc(x)
    t = torch.empty_like(xb)
    conv_nhwc(blk.body[0], xb, N, H, W, out_bf16=t, relu=True)
    out = torch.empty_like(xf)
    conv_nhwc(blk.body[2], t, N, H, W, residual=xf, out_f32=out, alpha=float(blk.res_scale))
    return ops.nhwc_to_nchw(out)


def _ca_params(ca):
    c0, c2 = ca.conv_du[0], ca.conv_du[2]
    Cr, C = c0.weight.shape[:2]
    return (c0.weight.detach().reshape(Cr, C).contiguous(), c0.bias.detach(),
            c2.weight.detach().reshape(C, Cr).contiguous(), c2.bias.detach())


def ca_forward(ca, x):
    """CALayer.forward (reference architectures.py:41-44): x * sigmoid(FC(relu(FC(avgpool(x)))))."""
    x = _require_cuda(x)
    N, C, H, W = x.shape
    xf, _ = ops.nchw_to_nhwc(x, want_bf16=False)
    pp = ops.pool_sum(xf)
    zero = torch.zeros_like(xf)
    out = torch.empty_like(xf)
    outb = torch.empty(xf.shape, dtype=torch.bfloat16, device=x.device)
    w1, b1, w2, b2 = _ca_params(ca)
    ops.ca_apply(pp, xf, zero, w1, b1, w2, b2, out, outb, N=N, H=H, W=W, C=C)
    return ops.nhwc_to_nchw(out)


def _rcab_nhwc(blk, xf, xb, N, H, W, C):
    t = torch.empty_like(xb)
    conv_nhwc(blk.body[0], xb, N, H, W, out_bf16=t, relu=True)
    u = torch.empty_like(xf)
    pp = torch.empty((N, ops.pool_rows(H, W, C), C), dtype=torch.float32, device=xf.device)
    conv_nhwc(blk.body[2], t, N, H, W, out_f32=u, pool_partial=pp)
    out = torch.empty_like(xf)
    outb = torch.empty_like(xb)
    w1, b1, w2, b2 = _ca_params(blk.body[3])
    ops.ca_apply(pp, u, xf, w1, b1, w2, b2, out, outb, N=N, H=H, W=W, C=C)
    return out, outb


def rcab_forward(blk, x):
    """RCAB.forward (reference architectures.py:81-84)."""
    x = _require_cuda(x)
    N, C, H, W = x.shape
    xf, xb = ops.nchw_to_nhwc(x)
    out, _ = _rcab_nhwc(blk, xf, xb, N, H, W, C)
    return ops.nhwc_to_nchw(out)


def resgroup_forward(grp, x):
    """ResidualGroup.forward (reference architectures.py:121-124)."""
    x = _require_cuda(x)
    N, C, H, W = x.shape
    xf, xb = ops.nchw_to_nhwc(x)
    cf, cb = xf, xb
    mods = list(grp.body)
    for blk in mods[:-1]:
        cf, cb = _rcab_nhwc(blk, cf, cb, N, H, W, C)
    out = torch.empty_like(xf)
    conv_nhwc(mods[-1], cb, N, H, W, residual=xf, out_f32=out)
    return ops.nhwc_to_nchw(out)


def upsampler_forward(up, x):
    """Upsampler.forward (reference common.py:29-44): conv + PixelShuffle stages, shuffle folded in the store."""
    x = _require_cuda(x)
    N, C, H, W = x.shape
    _, cur = ops.nchw_to_nhwc(x, want_f32=False)
    mods = list(up)
    for conv, ps in zip(mods[0::2], mods[1::2]):
        r = ps.upscale_factor
        out = torch.empty((N, H * r, W * r, C), dtype=torch.bfloat16, device=x.device)
        conv_nhwc(conv, cur, N, H, W, out_bf16=out, out_shuffle_r=r)
        cur, H, W = out, H * r, W * r
    return ops.nhwc_to_nchw(cur)
