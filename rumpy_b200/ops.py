"""Tensor-level wrappers over the C ABI (include/rumpy_b200.h).

PyTorch is used here only for device memory and streams: every wrapper passes raw `data_ptr()`s and the
current CUDA stream to the shared library.  Layout convention: activations NHWC (`[N,H,W,C]` contiguous),
bf16 operand tensors + fp32 residual-stream tensors.
"""
from __future__ import annotations

import torch

from . import _lib

RELU = 1
POOL = 32


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f'{name}: expected contiguous CUDA {dtype} tensor, got {t.dtype} cuda={t.is_cuda} '
                         f'contiguous={t.is_contiguous()}')


def device_check():
    _lib.call('rumpy_device_check')


def pool_rows(H, W, C):
    """Rows per image of a pool_partial buffer (rumpy_pool_rows)."""
    return _lib.load().rumpy_pool_rows(int(H), int(W), int(C))


def pack_conv3x3(w_oihw, rows_padded=0, shuffle_r=1, dgrad=False, out=None):
    """OIHW fp32 -> packed bf16 [9][rows][k]  (rumpy_pack_conv3x3)."""
    _chk(w_oihw, torch.float32, 'w')
    cout, cin = w_oihw.shape[0], w_oihw.shape[1]
    rows = cin if dgrad else max(rows_padded, cout)
    k = cout if dgrad else cin
    if out is None:
        out = torch.empty((9, rows, k), dtype=torch.bfloat16, device=w_oihw.device)
    _lib.call('rumpy_pack_conv3x3', w_oihw.data_ptr(), out.data_ptr(), cout, cin, rows_padded, shuffle_r,
              int(dgrad), _stream())
    return out


def pack_bias(b, rows_padded=0, shuffle_r=1, out=None):
    _chk(b, torch.float32, 'bias')
    cout = b.shape[0]
    rows = max(rows_padded, cout)
    if out is None:
        out = torch.empty((rows,), dtype=torch.float32, device=b.device)
    _lib.call('rumpy_pack_bias', b.data_ptr(), out.data_ptr(), cout, rows_padded, shuffle_r, _stream())
    return out


def conv3x3(x, w_packed, bias=None, *, residual=None, mask=None, out_bf16=None, out_f32=None, pool_partial=None,
            N, H, W, Cin, Cout, in_unshuffle_r=1, out_shuffle_r=1, relu=False, alpha=1.0):
    """Fused tensor-core conv (rumpy_conv3x3).  H, W are the conv's own spatial size."""
    _chk(x, torch.bfloat16, 'x'); _chk(w_packed, torch.bfloat16, 'w_packed'); _chk(bias, torch.float32, 'bias')
    _chk(residual, torch.float32, 'residual'); _chk(mask, torch.bfloat16, 'mask')
    _chk(out_bf16, torch.bfloat16, 'out_bf16'); _chk(out_f32, torch.float32, 'out_f32')
    _chk(pool_partial, torch.float32, 'pool_partial')
    flags = (RELU if relu else 0) | (POOL if pool_partial is not None else 0)
    _lib.call('rumpy_conv3x3', x.data_ptr(), w_packed.data_ptr(), _ptr(bias), _ptr(residual), _ptr(mask),
              _ptr(out_bf16), _ptr(out_f32), _ptr(pool_partial), N, H, W, Cin, Cout, in_unshuffle_r, out_shuffle_r,
              flags, float(alpha), _stream())


def conv3x3_tail(x, w_packed16, bias16, out_nchw, *, N, H, W, Cin, cout_real):
    _chk(x, torch.bfloat16, 'x'); _chk(w_packed16, torch.bfloat16, 'w'); _chk(bias16, torch.float32, 'bias')
    _chk(out_nchw, torch.float32, 'out')
    _lib.call('rumpy_conv3x3_tail', x.data_ptr(), w_packed16.data_ptr(), bias16.data_ptr(), out_nchw.data_ptr(),
              N, H, W, Cin, cout_real, _stream())


def head_conv(x_nchw, w_oihw, bias, out_f32, out_bf16):
    _chk(x_nchw, torch.float32, 'x'); _chk(w_oihw, torch.float32, 'w'); _chk(bias, torch.float32, 'bias')
    _chk(out_f32, torch.float32, 'out_f32'); _chk(out_bf16, torch.bfloat16, 'out_bf16')
    N, Cin, H, W = x_nchw.shape
    C = w_oihw.shape[0]
    _lib.call('rumpy_head_conv', x_nchw.data_ptr(), w_oihw.data_ptr(), bias.data_ptr(), out_f32.data_ptr(),
              out_bf16.data_ptr(), N, H, W, Cin, C, _stream())


def ca_apply(pool_partial, u, x_in, w1, b1, w2, b2, x_out, x_out_bf16, *, N, H, W, C, save=None):
    """save = (mean[N,C], hid[N,Cr], y[N,C]) fp32 tensors or None."""
    Cr = w1.shape[0]
    u_is_f32 = u.dtype == torch.float32
    for t, n in ((w1, 'w1'), (b1, 'b1'), (w2, 'w2'), (b2, 'b2'), (x_in, 'x_in'), (x_out, 'x_out')):
        _chk(t, torch.float32, n)
    _chk(x_out_bf16, torch.bfloat16, 'x_out_bf16')
    sm, sh, sy = save if save is not None else (None, None, None)
    _lib.call('rumpy_ca_apply', pool_partial.data_ptr(), u.data_ptr(), int(u_is_f32), x_in.data_ptr(),
              w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), x_out.data_ptr(), x_out_bf16.data_ptr(),
              _ptr(sm), _ptr(sh), _ptr(sy), N, H, W, C, Cr, _stream())


def nchw_to_nhwc(x, want_f32=True, want_bf16=True):
    """fp32 NCHW -> (fp32 NHWC | None, bf16 NHWC | None)."""
    _chk(x, torch.float32, 'x')
    N, C, H, W = x.shape
    yf = torch.empty((N, H, W, C), dtype=torch.float32, device=x.device) if want_f32 else None
    yb = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    _lib.call('rumpy_nchw_to_nhwc', x.data_ptr(), _ptr(yf), _ptr(yb), N, C, H, W, _stream())
    return yf, yb


def nhwc_to_nchw(x):
    """fp32 / bf16 NHWC -> fp32 NCHW."""
    if x.dtype not in (torch.float32, torch.bfloat16) or not x.is_cuda or not x.is_contiguous():
        raise ValueError('nhwc_to_nchw: expected contiguous CUDA fp32/bf16 tensor')
    N, H, W, C = x.shape
    y = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    _lib.call('rumpy_nhwc_to_nchw', x.data_ptr(), int(x.dtype == torch.bfloat16), y.data_ptr(), N, C, H, W,
              _stream())
    return y


def pool_sum(x_nhwc):
    _chk(x_nhwc, torch.float32, 'x')
    N, H, W, C = x_nhwc.shape
    pp = torch.empty((N, pool_rows(H, W, C), C), dtype=torch.float32, device=x_nhwc.device)
    _lib.call('rumpy_pool_sum', x_nhwc.data_ptr(), pp.data_ptr(), N, H, W, C, _stream())
    return pp


def conv3x3_wgrad(g, x, dw, *, N, H, W, Cin, Cout, g_unshuffle_r=1, alpha=1.0, accumulate=False):
    """dW (OIHW fp32) of the 3x3 conv from bf16 NHWC gradient / activation operands (rumpy_conv3x3_wgrad)."""
    _chk(g, torch.bfloat16, 'g'); _chk(x, torch.bfloat16, 'x'); _chk(dw, torch.float32, 'dw')
    nbytes = _lib.load().rumpy_conv3x3_wgrad_workspace(N, H, W, Cin, Cout)
    if nbytes < 0:
        raise _lib.RumpyB200Error('wgrad workspace query failed')
    ws = torch.empty(nbytes, dtype=torch.uint8, device=g.device)
    _lib.call('rumpy_conv3x3_wgrad', g.data_ptr(), x.data_ptr(), dw.data_ptr(), ws.data_ptr(), N, H, W, Cin, Cout,
              g_unshuffle_r, float(alpha), int(accumulate), _stream())
    return dw
