"""Data parallelism for the native trunk: one process per GPU (torchrun), identical replicas, gradient
all-reduce over NCCL / NVLink.  Replaces the reference's single-process nn.DataParallel
(base_architecture.py:70-77): there the batch is scattered / outputs gathered every step; here every rank
trains on its own LR/HR patch batch and only the 62 MB (RCAN) flat fp32 gradient crosses NVLink, in buckets.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_parameters(module, src=0):
    """Identical replicas at step 0 (the reference replicates the module every step instead)."""
    if not is_distributed():
        return
    for p in module.parameters():
        dist.broadcast(p.data, src=src)


class GradAllReduce:
    """Sums a flat gradient buffer across ranks in fixed-size buckets on a side stream, so the NCCL kernels of
    bucket k overlap whatever the compute stream enqueues next.  The 1/world factor is folded into Adam."""

    def __init__(self, params=None, bucket_bytes=16 << 20):
        self.world_size = dist.get_world_size() if is_distributed() else 1
        self.bucket_elems = max(1, bucket_bytes // 4)
        self._stream = None

    def __call__(self, flat):
        if self.world_size == 1:
            return flat
        if flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=flat.device)
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                for off in range(0, flat.numel(), self.bucket_elems):
                    dist.all_reduce(flat[off:off + self.bucket_elems], op=dist.ReduceOp.SUM)
            torch.cuda.current_stream().wait_stream(self._stream)
        else:  # gloo (CPU tests of the host logic)
            for off in range(0, flat.numel(), self.bucket_elems):
                dist.all_reduce(flat[off:off + self.bucket_elems], op=dist.ReduceOp.SUM)
        return flat


    def chunked(self, flat, chunks):
        """All-reduce `flat` range by range as the backward finishes them: chunks = [(event | None, lo, hi)] from
        TrunkEngine.backward_chunks(); the side stream waits for each event and reduces [lo, hi) while the compute
        stream is still producing the earlier ranges.  Call right after the backward was enqueued."""
        if self.world_size == 1:
            return flat
        covered = sorted((lo, hi) for _, lo, hi in chunks)
        if covered[0][0] != 0 or covered[-1][1] != flat.numel() or any(a[1] != b[0] for a, b in zip(covered, covered[1:])):
            raise ValueError('gradient chunks do not tile the flat buffer')
        if not flat.is_cuda:     # gloo (CPU tests of the host logic)
            for _, lo, hi in chunks:
                for off in range(lo, hi, self.bucket_elems):
                    dist.all_reduce(flat[off:min(hi, off + self.bucket_elems)], op=dist.ReduceOp.SUM)
            return flat
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=flat.device)
        with torch.cuda.stream(self._stream):
            for event, lo, hi in chunks:
                self._stream.wait_event(event)
                for off in range(lo, hi, self.bucket_elems):
                    dist.all_reduce(flat[off:min(hi, off + self.bucket_elems)], op=dist.ReduceOp.SUM)
        torch.cuda.current_stream().wait_stream(self._stream)
        return flat


def shard_round_robin(items, rank=None, world=None):
    """Inference sharding: whole images / frames round-robin over ranks, no collective."""
    if rank is None:
        rank = dist.get_rank() if is_distributed() else 0
    if world is None:
        world = dist.get_world_size() if is_distributed() else 1
    return list(items)[rank::world]


class FramesInFlight:
    """Whole-frame inference with `depth` frames in flight on ONE GPU (the per-GPU half of image-sharded inference,
    SURVEY 8e: ranks take frames round-robin, each rank runs its frames through this).

    Each in-flight frame gets its own engine (own activation workspace; parameters are shared, the packed bf16
    weights are per engine) and its own CUDA stream; results are identical to running the frames one by one.
    Why more than one: a large frame runs one kernel per layer, compute-bound convs alternating with the HBM-bound
    channel-attention pass, and within one frame the two cannot overlap (the pass needs the global pool of the conv
    before it); a second frame on another stream can fill the gaps.  Measured on 1080p RCAN frames (bench.py
    `frame_1080p`, which times both and reports the faster): with the final conv kernel of round 2 (which fills the
    shared-memory pipe of every SM by itself) two in flight are 139.6 ms per frame against 141.0 one at a time; with the
    earlier conv kernel it was 147.6 against 159.0, and with the fp32 pre-attention activation of round 1's plan the
    other way round, 174.8 against 168.7."""

    def __init__(self, net, depth=1):
        import torch
        from . import engine as _engine
        if depth < 1:
            raise ValueError('depth >= 1')
        arch, kw = net._engine_kwargs()
        params = list(net.parameters())
        self.net = net
        self.engines = [net.native_engine()] + [_engine.TrunkEngine(arch, params, **kw) for _ in range(depth - 1)]
        dev = params[0].device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]

    def run(self, frames, consume=None):
        """frames: iterable of N x C x H x W CUDA tensors.  Returns the outputs in order (or passes each to
        `consume(index, out)` on the frame's stream and returns nothing, to keep memory flat)."""
        import torch
        main = torch.cuda.current_stream()
        outs = []
        with torch.no_grad():
            for i, f in enumerate(frames):
                k = i % len(self.engines)
                st = self.streams[k]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    out = self.engines[k].forward(f)
                    f.record_stream(st)
                    if consume is not None:
                        consume(i, out)
                    else:
                        outs.append(out)
            for st in self.streams:
                main.wait_stream(st)
        return None if consume is not None else outs
