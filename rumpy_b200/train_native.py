"""Native train step: forward -> L1 loss + gradient -> backward (tensor-core dgrad / wgrad) -> [all-reduce] ->
[grad-norm clip] -> fused Adam.  Restates BaseModel.run_train + standard_update of the reference
(base_architecture.py:425-440, 457-485) with every arithmetic step in librumpy_b200.so."""
from __future__ import annotations

import torch

from . import _lib

_scratch = {}


def _buf(key, shape, device, dtype=torch.float32):
    t = _scratch.get(key)
    if t is None or t.shape != torch.Size(shape) or t.device != device:
        t = torch.empty(shape, dtype=dtype, device=device)
        _scratch[key] = t
    return t


def l1_loss(out, y, want_grad=False, gscale=1.0):
    """nn.L1Loss() (mean).  Returns loss (0-dim device tensor) [and dy = gscale * sign(out - y) / numel]."""
    if not out.is_cuda:
        raise _lib.RumpyB200Error('rumpy_b200 l1_loss: CUDA tensors only (no CPU fallback)')
    out = out.contiguous().float()
    y = y.contiguous().float()
    if out.shape != y.shape:
        raise ValueError(f'l1_loss: shape mismatch {tuple(out.shape)} vs {tuple(y.shape)}')
    lib = _lib.load()
    ws = _buf(('l1ws', out.device), (lib.rumpy_l1_workspace_floats(),), out.device)
    loss = torch.empty((), dtype=torch.float32, device=out.device)
    dy = _buf(('dy', out.device), tuple(out.shape), out.device) if want_grad else None
    _lib.call('rumpy_l1_loss_grad', out.data_ptr(), y.data_ptr(), 0 if dy is None else dy.data_ptr(),
              loss.data_ptr(), ws.data_ptr(), out.numel(), float(gscale), torch.cuda.current_stream().cuda_stream)
    return (loss, dy) if want_grad else loss


def train_step(net, optimizer, x, y, grad_clip=None, allreduce=None, metadata=None, y_ready=None, after_loss=None):
    """One optimiser step on batch (x, y); returns (loss 0-dim device tensor, SR output on device).
    metadata: the [N, M, 1, 1] vector of the meta-attention networks (QRCAN), else None.
    y_ready: CUDA event after which `y` is valid (its host-to-device copy runs on another stream while the forward
    computes; the compute stream waits for it right before the loss).  after_loss(loss, out): called once the forward
    and the loss are enqueued -- the handler starts the device-to-host copy of the SR batch there, so that it
    overlaps the backward and the optimiser step."""
    eng = net.native_engine()
    if metadata is not None:
        eng.set_metadata(metadata, x.shape[0])
    if getattr(optimizer, 'flat_g', None) is not None and eng.flat_grads is not optimizer.flat_g:
        optimizer.attach_engine(eng)       # engine writes gradients straight into the optimiser's flat buffer
    out = eng.forward(x, training=True)
    if y_ready is not None:
        torch.cuda.current_stream().wait_event(y_ready)
    loss, dy = l1_loss(out, y, want_grad=True)
    if after_loss is not None:
        after_loss(loss, out)
    chunks = eng.backward_chunks() if (allreduce is not None and allreduce.world_size > 1) else None
    eng.backward(x, dy)
    flat_g = eng.flat_grads
    if allreduce is not None:
        # NCCL sum over ranks, range by range while the remaining weight-gradient chunks are still running;
        # 1/world folded into the Adam grad scale
        if chunks is not None:
            allreduce.chunked(flat_g, chunks)
        else:
            allreduce(flat_g)
        optimizer.grad_scale = 1.0 / allreduce.world_size
    if grad_clip is not None:
        coef = _buf(('clip', flat_g.device), (2,), flat_g.device)
        ws = _buf(('clipws', flat_g.device), (1024,), flat_g.device)
        # clip_grad_norm_ acts on the averaged gradient: ||g/world|| = ||g|| / world
        _lib.call('rumpy_grad_clip_coef', flat_g.data_ptr(), flat_g.numel(),
                  float(grad_clip) / float(optimizer.grad_scale if allreduce is not None else 1.0), coef.data_ptr(),
                  ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        optimizer.grad_scale_dev = coef
    if getattr(optimizer, 'flat_g', None) is not flat_g:
        # foreign optimiser (e.g. torch's RMSprop from the reference's optimizer_type switch): it knows nothing of
        # the fused step's grad scales, so the 1/world average and the clip coefficient are applied to the flat
        # gradient here, then the gradients are handed over as .grad
        if allreduce is not None and allreduce.world_size > 1:
            flat_g.mul_(1.0 / allreduce.world_size)
        if grad_clip is not None:
            flat_g.mul_(optimizer.grad_scale_dev[0])
        optimizer.grad_scale, optimizer.grad_scale_dev = 1.0, None
        for p, g in zip(eng.params, eng.grad_views()):
            p.grad = g
        optimizer.step()
        eng.invalidate()                   # parameters changed behind the engine's back: repack before the next use
        return loss, out
    optimizer.step()
    return loss, out
