"""Whole-network executor binding (rumpy_net_* in include/rumpy_b200.h).

`TrunkEngine` owns, for one RCAN / EDSR module: the native net handle, the packed bf16 weight buffer, the
activation workspace and (optionally) a captured CUDA graph of the forward.  PyTorch supplies device
memory and streams only; every FLOP runs in librumpy_b200.so.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

ARCH_RCAN, ARCH_EDSR, ARCH_QRCAN, ARCH_QEDSR, ARCH_HAN, ARCH_QHAN = 0, 1, 2, 3, 4, 5

PARAM_EPOCH = [0]   # bumped by every native optimiser step (FusedAdam writes parameters through the flat buffer, which
                    # leaves the per-parameter version counters alone): part of the stand-alone blocks' cache keys

_FLAT = {}   # id(first parameter) -> (flat fp32 buffer, weakref to first parameter): shared by engine and FusedAdam


def flatten_parameters(params):
    """Makes every parameter a view of ONE flat fp32 buffer (state_dict / load_state_dict keep working: they copy
    in place).  Gives O(1) change detection (views share the base tensor's version counter), one-kernel Adam and a
    single all-reduce buffer.  Idempotent."""
    import weakref
    params = list(params)
    ent = _FLAT.get(id(params[0]))
    if ent is not None and ent[1]() is params[0]:
        flat, off, ok = ent[0], 0, True
        for p in params:
            if p.data_ptr() != flat.data_ptr() + off * 4:
                ok = False
                break
            off += p.numel()
        if ok and off == flat.numel():
            return flat
    dev = params[0].device
    n = sum(p.numel() for p in params)
    flat = torch.empty(n, dtype=torch.float32, device=dev)
    off = 0
    with torch.no_grad():
        for p in params:
            k = p.numel()
            flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat[off:off + k].view(p.shape)
            off += k
    _FLAT[id(params[0])] = (flat, weakref.ref(params[0]))
    return flat


class TrunkEngine:
    def __init__(self, arch, params, *, n_feats, n_groups, n_blocks, reduction=16, scale=4, res_scale=1.0,
                 in_feats=3, out_feats=3, u_f32=True, num_metadata=0, q_hidden=0, block_has_q=None, modulate=False,
                 q_relu=True):
        self.lib = _lib.load()
        self.arch, self.scale, self.in_feats, self.out_feats = arch, scale, in_feats, out_feats
        self.params = list(params)
        h = ctypes.c_void_p()
        self._meta = None
        if arch in (ARCH_QRCAN, ARCH_QEDSR, ARCH_QHAN):
            flags = bytes(bytearray(int(bool(f)) for f in block_has_q))
            if len(flags) != (n_blocks if arch == ARCH_QEDSR else n_groups * n_blocks):
                raise ValueError('block_has_q needs one flag per RCAB / ResBlock')
            _lib.call('rumpy_net_create_q', ctypes.byref(h), {ARCH_QRCAN: 0, ARCH_QEDSR: 1, ARCH_QHAN: 2}[arch], n_feats, n_groups, n_blocks, reduction,
                      scale, float(res_scale), in_feats, out_feats, int(num_metadata), int(q_hidden), flags,
                      int(bool(modulate)), int(bool(q_relu)))
        else:
            _lib.call('rumpy_net_create', ctypes.byref(h), 2 if arch == ARCH_HAN else arch, n_feats, n_groups, n_blocks, reduction, scale,
                      float(res_scale), in_feats, out_feats, int(u_f32))
        self.handle = h
        for name, value in _lib.env_options().items():
            _lib.call('rumpy_net_set_option', h, name.encode(), int(value))
        n = self.lib.rumpy_net_num_params(h)
        if n != len(self.params):
            raise _lib.RumpyB200Error(f'parameter count mismatch: native {n} vs module {len(self.params)}')
        self.device = self.params[0].device
        if self.device.type != 'cuda':
            raise _lib.RumpyB200Error('rumpy_b200 has no CPU path: move the model to a CUDA (sm_100) device')
        self.packed = torch.empty(self.lib.rumpy_net_packed_bytes(h, 1), dtype=torch.uint8, device=self.device)
        self._packed_training = False
        self.flat_params = flatten_parameters(self.params)   # all parameters are views of one flat buffer
        self._flat_sig = None
        self.flat_grads = None
        self._grad_views = None
        self._grad_ptr_array = None
        self._ptr_array = None
        self._ptr_sig = None
        self._pack_sig = None
        self._ws = {}
        self._graphs = {}

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.rumpy_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ per-handle options
    def set_option(self, name, value):
        """rumpy_net_set_option: execution knobs of THIS engine ('trunk', 'cluster', 'cluster_groups', 'band',
        'trunk_bwd', 'fused_ca', 'wgrad_chunks', ...).  Plans, workspaces and graphs built under the old value go."""
        _lib.call('rumpy_net_set_option', self.handle, name.encode(), int(value))
        self._ws.clear()
        self._graphs.clear()
        self._last_infer_shape = None
        self._chunk_key = None

    def get_option(self, name):
        return int(self.lib.rumpy_net_get_option(self.handle, name.encode()))

    def set_trunk_events(self, ev0, ev1):
        """CUDA events recorded right before / after the trunk kernel of every forward (None, None = off)."""
        _lib.call('rumpy_net_set_trunk_events', self.handle, None if ev0 is None else ev0.cuda_event,
                  None if ev1 is None else ev1.cuda_event)
        self._graphs.clear()
        self._last_infer_shape = None

    def set_timeline(self, buf, layers=0):
        """Device int64 tensor the kernels fill with clock64 stamps (None = off)."""
        _lib.call('rumpy_net_set_timeline', self.handle, None if buf is None else buf.data_ptr(), int(layers))
        self._timeline = buf
        self._graphs.clear()
        self._last_infer_shape = None

    # ------------------------------------------------------------------ parameter tracking
    def _param_ptrs(self):
        if self.flat_params is not None and self._ptr_sig is not None and \
                self._flat_sig == (self.flat_params.data_ptr(), self.params[0].data_ptr(), self.params[-1].data_ptr()):
            return self._ptr_array            # O(1): parameters are views of one flat buffer
        sig = tuple(p.data_ptr() for p in self.params)
        if sig != self._ptr_sig:
            for p in self.params:
                if p.dtype != torch.float32 or not p.is_contiguous() or p.device != self.device:
                    raise _lib.RumpyB200Error('parameters must be contiguous fp32 tensors on one CUDA device')
            self._ptr_array = (ctypes.c_void_p * len(sig))(*sig)
            self._ptr_sig = sig
            self._pack_sig = None
            self._graphs.clear()
        if self.flat_params is not None:
            self._flat_sig = (self.flat_params.data_ptr(), self.params[0].data_ptr(), self.params[-1].data_ptr())
        return self._ptr_array

    def attach_flat(self, flat_params, flat_grads):
        """Called by FusedAdam: parameters (and .grad) are views into these flat fp32 buffers."""
        if flat_params.data_ptr() != self.flat_params.data_ptr():
            raise _lib.RumpyB200Error('optimizer and engine disagree on the flat parameter buffer')
        self.flat_params, self.flat_grads = flat_params, flat_grads
        self._ptr_sig = None
        self._flat_sig = None
        self._grad_views = None

    def _version_sig(self):
        # FusedAdam bumps the flat buffer's counter (one in-place op per step); load_state_dict, foreign optimisers
        # and user code write through the parameters, whose counters are their own (`p.data = view` does not share
        # the base tensor's counter) -- so both are part of the signature.
        flat_v = self.flat_params._version if self.flat_params is not None else 0
        return (flat_v, sum(p._version for p in self.params))

    def invalidate(self):
        """Forget everything derived from the parameter VALUES or from buffer addresses: packed bf16 weights are
        rebuilt on the next forward, captured graphs are dropped."""
        self._pack_sig = None
        self._graphs.clear()
        self._last_infer_shape = None

    def refresh_weights(self, force=False, training=False):
        """Repacks fp32 OIHW parameters into the bf16 tensor-core operand layout when they changed."""
        ptrs = self._param_ptrs()
        sig = self._version_sig()
        if force or sig != self._pack_sig or (training and not self._packed_training):
            _lib.call('rumpy_net_pack', self.handle, ptrs, self.packed.data_ptr(), int(training),
                      torch.cuda.current_stream().cuda_stream)
            self._pack_sig = sig
            self._packed_training = bool(training)
        return ptrs

    def workspace(self, N, H, W, training):
        key = (N, H, W, bool(training))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = self.lib.rumpy_net_workspace_bytes(self.handle, N, H, W, int(training))
            if nbytes < 0:
                msg = self.lib.rumpy_last_error()
                raise _lib.RumpyB200Error('workspace query failed: ' + (msg.decode() if msg else '?'))
            if any(k[3] == key[3] for k in self._ws):
                # one workspace per mode: the other shape's buffer is released, and with it every CUDA graph that
                # was captured against its address (replaying one would touch freed memory)
                self._ws = {k: v for k, v in self._ws.items() if k[3] != key[3]}
                self._graphs.clear()
                self._last_infer_shape = None
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ Q-RCAN metadata
    def set_metadata(self, metadata, N):
        """metadata: [N, M, 1, 1] or [N, M] tensor (what QRCAN.forward receives).  It is copied into a buffer the
        engine owns (static address: CUDA-graph replays see the new values)."""
        if self.arch not in (ARCH_QRCAN, ARCH_QEDSR, ARCH_QHAN):
            raise _lib.RumpyB200Error('set_metadata: not a meta-attention engine')
        if metadata is None:
            raise RuntimeError('Metadata needs to be specified for this network to run properly.')
        m = metadata.reshape(metadata.shape[0], -1).to(device=self.device, dtype=torch.float32)
        if m.shape[0] != N:
            raise ValueError(f'metadata for {m.shape[0]} images, batch has {N}')
        if self._meta is None or self._meta.shape != m.shape:
            self._meta = torch.empty_like(m)
            self._graphs.clear()
        self._meta.copy_(m)
        _lib.call('rumpy_net_set_metadata', self.handle, self._meta.data_ptr(), int(N), int(m.shape[1]))

    # ------------------------------------------------------------------ forward
    def _check_input(self, x):
        if x.device != self.device:
            raise _lib.RumpyB200Error(f'input on {x.device}, model on {self.device} (no CPU fallback)')
        if x.dim() != 4 or x.shape[1] != self.in_feats:
            raise ValueError(f'expected N x {self.in_feats} x H x W input, got {tuple(x.shape)}')
        return x.contiguous().float()

    def forward(self, x, training=False, out=None):
        x = self._check_input(x)
        N, _, H, W = x.shape
        ptrs = self.refresh_weights(training=training)
        ws = self.workspace(N, H, W, training)
        if out is None:
            out = torch.empty((N, self.out_feats, H * self.scale, W * self.scale), dtype=torch.float32,
                              device=self.device)
        _lib.call('rumpy_net_forward', self.handle, ptrs, self.packed.data_ptr(), x.data_ptr(), out.data_ptr(),
                  ws.data_ptr(), N, H, W, int(training), torch.cuda.current_stream().cuda_stream)
        return out

    # ------------------------------------------------------------------ backward (after forward(training=True))
    def grad_views(self):
        """Per-parameter gradient tensors (views of one flat fp32 buffer), state_dict order."""
        if self._grad_views is None:
            if self.flat_grads is None:
                self.flat_grads = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32,
                                              device=self.device)
            views, off = [], 0
            for p in self.params:
                views.append(self.flat_grads[off:off + p.numel()].view(p.shape))
                off += p.numel()
            self._grad_views = views
            self._grad_ptr_array = (ctypes.c_void_p * len(views))(*[v.data_ptr() for v in views])
        return self._grad_views

    def backward(self, x, dy):
        """d(loss)/d(params) for the last forward(x, training=True); dy = upstream gradient (fp32 NCHW)."""
        x = self._check_input(x)
        N, _, H, W = x.shape
        dy = dy.contiguous().float()
        if tuple(dy.shape) != (N, self.out_feats, H * self.scale, W * self.scale) or dy.device != self.device:
            raise ValueError(f'bad upstream gradient {tuple(dy.shape)} on {dy.device}')
        grads = self.grad_views()
        ptrs = self._param_ptrs()
        ws = self.workspace(N, H, W, True)
        _lib.call('rumpy_net_backward', self.handle, ptrs, self.packed.data_ptr(), x.data_ptr(), dy.data_ptr(),
                  self._grad_ptr_array, ws.data_ptr(), N, H, W, torch.cuda.current_stream().cuda_stream)
        return grads

    # ------------------------------------------------------------------ chunked backward (data parallel)
    def backward_chunks(self):
        """[(event, lo, hi)]: after `event` (recorded by the native backward on its stream) the flat gradient range
        [lo, hi) is final.  Ranges run from the end of the buffer to its start and cover it exactly once."""
        first = (ctypes.c_int * 8)()
        k = self.lib.rumpy_net_backward_chunks(self.handle, first, 8)
        if k <= 0:
            raise _lib.RumpyB200Error('backward_chunks: no training plan yet (run forward(training=True) first)')
        key = tuple(first[i] for i in range(k))
        if getattr(self, '_chunk_key', None) != key:
            offs = [0]
            for p in self.params:
                offs.append(offs[-1] + p.numel())
            events = [torch.cuda.Event() for _ in range(k)]
            for e in events:
                e.record()                       # materialises the cudaEvent_t handle
            arr = (ctypes.c_void_p * k)(*[e.cuda_event for e in events])
            _lib.call('rumpy_net_set_backward_events', self.handle, arr, k)
            hi, chunks = offs[-1], []
            for i in range(k):
                lo = offs[first[i]]
                chunks.append((events[i], lo, hi))
                hi = lo
            self._chunk_key, self._chunks, self._chunk_events_arr = key, chunks, arr
        return self._chunks

    def forward_inference(self, x):
        """Module-level inference entry: the first call with a shape launches eagerly; repeated calls with the same
        shape replay a captured CUDA graph (launch overhead of ~600 kernels -> one graph launch)."""
        key = tuple(x.shape)
        if getattr(self, '_last_infer_shape', None) == key:
            return self.forward_graphed(x).clone()
        self._last_infer_shape = key
        return self.forward(x)

    # ------------------------------------------------------------------ CUDA-graph replay (inference)
    def forward_graphed(self, x):
        """Forward through a captured CUDA graph (static shapes, weights assumed unchanged between calls
        unless refresh_weights() sees new versions -- repacking happens outside the graph)."""
        x = self._check_input(x)
        self.refresh_weights()
        key = tuple(x.shape)
        g = self._graphs.get(key)
        if g is not None and g[3].data_ptr() != self.workspace(x.shape[0], x.shape[2], x.shape[3], False).data_ptr():
            g = None                               # captured against a workspace that has been replaced
        if g is None:
            sx = torch.empty_like(x)
            sy = torch.empty((x.shape[0], self.out_feats, x.shape[2] * self.scale, x.shape[3] * self.scale),
                             dtype=torch.float32, device=self.device)
            sx.copy_(x)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.forward(sx, out=sy)      # warm-up: builds the plan, sets kernel attributes
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.forward(sx, out=sy)
            # the graph keeps the workspace it was captured with alive
            g = (graph, sx, sy, self.workspace(x.shape[0], x.shape[2], x.shape[3], False))
            self._graphs = {key: g}
        graph, sx, sy, _ = g
        sx.copy_(x)
        graph.replay()
        return sy
