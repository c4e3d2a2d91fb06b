"""Fused Adam for the native trunk (reference: optim.Adam at base_architecture.py:93-95, stepped at :437).

A `torch.optim.Optimizer` subclass so `param_groups` (schedulers write `lr` there), `state_dict()` /
`load_state_dict()` keep the reference's checkpoint format (`'optimizer'` entry, index-keyed Adam state with
`step`, `exp_avg`, `exp_avg_sq`).  The update itself is ONE kernel over flat fp32 buffers
(rumpy_adam_step): parameters, gradients and both moments are views into four flat tensors.
"""
from __future__ import annotations

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError('FusedAdam: a single parameter group is supported')
        self._params = list(self.param_groups[0]['params'])
        if not self._params or any(p.device.type != 'cuda' or p.dtype != torch.float32 for p in self._params):
            raise _lib.RumpyB200Error('FusedAdam needs fp32 CUDA parameters (no CPU fallback)')
        from .engine import flatten_parameters
        dev = self._params[0].device
        self.flat_p = flatten_parameters(self._params)               # parameters become views of flat_p
        n = self.flat_p.numel()
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self._step = 0
        off = 0
        self._slices = []
        for p in self._params:
            k = p.numel()
            p.grad = self.flat_g[off:off + k].view(p.shape)
            self._slices.append((off, k))
            off += k
        self.grad_scale_dev = None       # optional device float (clip coefficient), consumed by the next step
        self.grad_scale = 1.0
        self._engines = []
        self._sync_state()

    def attach_engine(self, engine):
        engine.attach_flat(self.flat_p, self.flat_g)
        self._engines.append(engine)

    def _sync_state(self):
        """Exposes torch.optim.Adam-shaped per-parameter state (views of the flat moment buffers)."""
        for p, (off, k) in zip(self._params, self._slices):
            self.state[p] = {'step': torch.tensor(float(self._step)),
                             'exp_avg': self.flat_m[off:off + k].view(p.shape),
                             'exp_avg_sq': self.flat_v[off:off + k].view(p.shape)}

    def zero_grad(self, set_to_none=False):
        """Clears the flat gradient buffer (p.grad are views of it, so `set_to_none` is ignored).  The native train
        step does not call this -- its backward overwrites every gradient -- but the module-level autograd path
        (`net(x)`, `loss.backward()`, `opt.step()`) accumulates into .grad like any torch optimiser expects."""
        self.flat_g.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        self._step += 1
        _lib.call('rumpy_adam_step', self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                  self.flat_v.data_ptr(), self.flat_p.numel(), float(g['lr']), float(g['betas'][0]),
                  float(g['betas'][1]), float(g['eps']), self._step,
                  0 if self.grad_scale_dev is None else self.grad_scale_dev.data_ptr(), float(self.grad_scale),
                  torch.cuda.current_stream().cuda_stream)
        self.flat_p[:1].add_(0)  # bump the flat buffer's version counter -> engines repack
        from . import engine
        engine.PARAM_EPOCH[0] += 1   # ... and the stand-alone blocks' packed-weight caches (blocks_native._packed)
        self.grad_scale_dev = None
        self.grad_scale = 1.0
        for st in self.state.values():
            st['step'] = torch.tensor(float(self._step))

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = []
        for p, (off, k) in zip(self._params, self._slices):
            st = self.state.get(p, {})
            if 'exp_avg' in st:
                self.flat_m[off:off + k].copy_(st['exp_avg'].reshape(-1))
                self.flat_v[off:off + k].copy_(st['exp_avg_sq'].reshape(-1))
                steps.append(int(float(st['step'])))
        self._step = max(steps) if steps else 0
        self._sync_state()
