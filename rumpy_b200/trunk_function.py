"""Autograd boundary of the native trunk: one Function for the whole network (no per-layer Python)."""
from __future__ import annotations

import torch


class _TrunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module.native_engine()
        ctx.eng = eng
        ctx.x_shape = tuple(x.shape)
        ctx.save_for_backward(x)
        return eng.forward(x, training=True)

    @staticmethod
    def backward(ctx, gy):
        eng = ctx.eng
        (x,) = ctx.saved_tensors
        grads = eng.backward(x, gy.contiguous())
        return (None, None) + tuple(grads)


def trunk_apply(module, x):
    needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())
    if not needs_grad:
        return module.native_engine().forward_inference(x)
    return _TrunkFn.apply(module, x, *module.parameters())
