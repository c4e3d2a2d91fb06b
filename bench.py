"""Benchmark of the hot path.  Headline (BASELINE.json configs[1]): RCAN x4 (10 groups x 20 RCAB, 64 ch) forward on
synthetic 48x48 LR patches, batch 16 per GPU; metric = output Mpix/s.

    python bench.py --gpus 1 --steps 50 --warmup 5                 # this repo's sm_100a path
    python bench.py --impl reference --steps 5 --warmup 1          # the reference's own CPU path (baseline/_ref)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N ...                                      # weak scaling, one rank per GPU
    ... bench.py --gpus N --metric train                           # headline = the data-parallel train step (the leg
                                                                   # with a collective), same JSON contract

One JSON line on stdout (rank 0):
  value         whole-job throughput, inputs resident in HBM (CUDA-event device time per step, L2 flushed between
                steps, max over ranks)
  e2e           the same metric through the reference-facing handler call with pinned HOST buffers in and out
  roofline      the dominant kernel (the trunk kernel: 411 fused 64->64 convs) timed live with CUDA events recorded
                around it on the launching stream, against MEASURED_PEAKS.json
  cpu_baseline  the reference's CPU implementation (baseline/_ref when installed, else the oracle port) on the host
  train / edsr_full_train / frame_1080p   BASELINE configs[2] / [3] / [4] at the SAME N: data-parallel training with the
                NCCL gradient all-reduce (RCAN, EDSR-full) and whole 1080p frames sharded round-robin
  dp_check      (N > 1) the all-reduced gradient equals the mean of the ranks' own gradients
Everything else (Q-RCAN / HAN lines, eager-GPU baselines, glue kernels) goes to stderr and gpurun_out/bench_extra.json.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import recipe  # noqa: E402

WORKLOAD = 'RCAN x4 (10 groups x 20 RCAB, 64 ch) forward, synthetic 48x48 LR patches, batch 16 per GPU'
TRAIN_WORKLOAD = ('RCAN x4 (10 groups x 20 RCAB, 64 ch) train step (fwd + L1 + bwd + Adam 1e-4), synthetic 64x64 LR '
                  'patches, batch 16 per GPU, data parallel')
BATCH, LR_HW, SCALE, CH = 16, 48, 4, 64
TRAIN_BATCH, TRAIN_HW = 16, 64
FLOP_PER_LR_PIXEL = 31835520          # 2*MAC over all convs of RCAN x4
EDSR_FULL_FLOP_PER_LR_PIXEL = 100505088
CONV64_FLOP_PER_PIXEL = 2 * 64 * 64 * 9
OUT_MPIX_PER_STEP = BATCH * (LR_HW * SCALE) ** 2 / 1e6

_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=p['bf16_tflops'], tflops_sustained=p['bf16_tflops_sustained'], hbm=p['hbm_gbs'],
                    source='MEASURED_PEAKS.json (measured)')
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0, source='B200_PROFILING.md fallback')


def ev():
    return torch.cuda.Event(enable_timing=True)


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x2: 'applications_clocks_setting', 0x10: 'sync_boost',
               0x100: 'display_clock_setting'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.1)    # 10 Hz: NVML queries contend with the CUDA driver lock, 50 Hz slowed the host-side e2e loop

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable']}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------- workload
def make_state_dict():
    return recipe.make_weights(recipe.rcan_spec(10, 20, CH, 16, SCALE), seed=8)


# ------------------------------------------------------------------------------------------- the reference on the CPU
def reference_cpu_handler(threads, eval_mode=True):
    """The reference's own implementation of the path on the host cores.  kind 'reference': the UNMODIFIED reference
    from baseline/_ref (tools/install_reference.py) through ITS public API -- `define_model('rcan', ...)` ->
    `RCANHandler.run_eval / run_train` (rumpy/SISR/models/advanced/handlers.py:25-42, base_architecture.py:457-520);
    kind 'port': the oracle's torch-CPU restatement (oracle/sr_torch_cpu.py) when the install is absent."""
    torch.set_num_threads(threads)
    sd = {k: torch.from_numpy(v) for k, v in make_state_dict().items()}
    from oracle import ref_import
    if os.path.isdir(os.path.join(ref_import.INSTALLED, 'rumpy')):
        import contextlib
        ref_import.import_reference(ref_import.INSTALLED)
        with contextlib.redirect_stdout(sys.stderr):
            from rumpy.shared_framework.models import define_model
            h = define_model('rcan', device=torch.device('cpu'), model_save_dir=tempfile.mkdtemp(), eval_mode=eval_mode,
                             scale=SCALE, lr=1e-4)
        h.net.load_state_dict(sd, strict=True)

        def fwd(x):
            return h.run_eval(x)[0]

        def step(x, y):
            return h.run_train(x, y)[0]
        return 'reference', fwd, step
    from oracle import sr_torch_cpu

    def fwd(x):
        with torch.no_grad():
            return sr_torch_cpu.rcan_forward(sd, x, 10, 20, SCALE)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)

    def step(x, y):
        loss = torch.nn.functional.l1_loss(sr_torch_cpu.rcan_forward(params, x, 10, 20, SCALE), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss.item()
    return 'port', fwd, step


def time_cpu(fwd, batch, min_seconds, max_iters):
    x = torch.from_numpy(recipe.make_input((batch, 3, LR_HW, LR_HW), seed=8))
    fwd(x)  # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < max_iters and (time.perf_counter() - t_start < min_seconds or len(times) < 2):
        t0 = time.perf_counter()
        fwd(x)
        times.append(time.perf_counter() - t0)
    return float(np.median(times)), len(times)


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path, all host threads, rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    train = args.metric == 'train'
    kind, fwd, step = reference_cpu_handler(cores, eval_mode=not train)
    full, hw = (TRAIN_BATCH, TRAIN_HW) if train else (BATCH, LR_HW)
    xs = torch.from_numpy(recipe.make_input((full, 3, hw, hw), seed=8))
    ys = torch.from_numpy(recipe.make_input((full, 3, hw * SCALE, hw * SCALE), seed=9))
    # bound the whole run to a few minutes: one probe patch gives the per-patch cost, the per-step sample shrinks to fit
    t0 = time.perf_counter()
    step(xs[:1], ys[:1]) if train else fwd(xs[:1])
    per_patch = time.perf_counter() - t0
    budget, batch = 150.0, full
    while batch > 1 and per_patch * batch * (args.steps + args.warmup) > budget:
        batch //= 2
    x, y = xs[:batch].contiguous(), ys[:batch].contiguous()
    run = (lambda: step(x, y)) if train else (lambda: fwd(x))
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = (time.perf_counter() - t0) / args.steps
    if train:
        metric, unit, value, workload = 'RCAN x4 train patches/s', 'patches/s', batch / dt, TRAIN_WORKLOAD
    else:
        metric, unit, value, workload = 'RCAN x4 output Mpix/s (infer)', 'Mpix/s', batch * (hw * SCALE) ** 2 / 1e6 / dt, WORKLOAD
    impl = 'unmodified reference from baseline/_ref through its own handler API' if kind == 'reference' else \
        'oracle port (baseline/_ref not installed)'
    sample = f'{batch} of {full} patches per step x {args.steps} steps, fp32, torch-CPU (oneDNN), {impl}'
    emit({
        'impl': 'reference', 'metric': metric, 'value': value, 'unit': unit,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3 * (full / batch),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload, 'batch_per_gpu': full, 'lr_patch': hw, 'scale': SCALE},
        'cpu_baseline': {'value': value, 'unit': unit, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    })


# ------------------------------------------------------------------------------------------- rooflines
def conv_kernel_roofline(device, pk):
    """The per-layer conv kernel alone: tcgen05 conv 64->64 (+bias+ReLU) at the workload's activation shape, 20 launches
    in a CUDA graph, timed with CUDA events on the launching stream (the kernel shapes outside the trunk kernels use)."""
    from rumpy_b200 import ops
    N, H, W, C = BATCH, LR_HW, LR_HW, CH
    x = torch.rand((N, H, W, C), device=device).to(torch.bfloat16)
    w = (torch.rand((C, C, 3, 3), device=device) - 0.5) / 24
    b = torch.rand((C,), device=device)
    wp = ops.pack_conv3x3(w)
    y = torch.empty_like(x)
    reps = 20
    s = torch.cuda.Stream(device=device)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            ops.conv3x3(x, wp, b, out_bf16=y, N=N, H=H, W=W, Cin=C, Cout=C, relu=True)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            ops.conv3x3(x, wp, b, out_bf16=y, N=N, H=H, W=W, Cin=C, Cout=C, relu=True)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    best = None
    for _ in range(5):
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / reps
        best = t if best is None else min(best, t)
    flops = CONV64_FLOP_PER_PIXEL * N * H * W
    achieved = flops / best * 1e-12
    return {'kernel': 'conv3x3_tc_kernel<64,resident-B> (64->64, bias+ReLU, 16x48x48)', 'achieved': achieved,
            'frac': achieved / pk['tflops'], 'us_per_launch': best * 1e6}


def trunk_kernel_roofline(eng, x_dev, flush, pk):
    """Dominant kernel of the headline forward: the trunk kernel that runs all 411 body convs (64->64, 3x3) with
    their epilogues.  Timed live with CUDA events the library records right before / after the kernel on the launching
    stream (per-handle hook rumpy_net_set_trunk_events; eager forwards, L2 flushed before each): average and best of
    the timed launches."""
    mode = eng.lib.rumpy_net_trunk_mode(eng.handle)
    if mode <= 0:
        return None
    e0, e1 = ev(), ev()
    e0.record(); e1.record()
    torch.cuda.synchronize()
    eng.set_trunk_events(e0, e1)
    times = []
    try:
        with torch.no_grad():
            for i in range(12):
                flush.zero_()
                eng.forward(x_dev)
                torch.cuda.synchronize()
                if i >= 2:
                    times.append(e0.elapsed_time(e1) * 1e-3)
    finally:
        eng.set_trunk_events(None, None)
    avg = float(np.mean(times))
    n_convs = 10 * (2 * 20 + 1) + 1
    flops = n_convs * CONV64_FLOP_PER_PIXEL * BATCH * LR_HW * LR_HW
    achieved = flops / avg * 1e-12
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get('trunk_dram_bytes_per_launch')
    name = {1: 'trunk_pipe_kernel (persistent dataflow, epoch flags)',
            2: 'trunk_cluster_kernel (one thread-block cluster per image, DSMEM halo exchange)',
            3: 'trunk_band_kernel (role-swapped: weights in TMEM, N = 144 pixels per MMA)'}[mode]
    return {'bound': 'tensor', 'kernel': name + f': {n_convs} fused conv3x3 64->64 layers + CA + skips, 16x48x48',
            'achieved': achieved, 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'frac': achieved / pk['tflops'],
            'frac_of_sustained_peak': achieved / pk['tflops_sustained'],
            'traffic': traffic, 'us_per_launch': avg * 1e6, 'us_per_launch_best': float(min(times)) * 1e6,
            'launches_timed': len(times), 'flops_per_launch': flops,
            'peak_source': pk['source'] + ', burst figure (one ~2.5 ms kernel timed alone, L2 flushed)'}


# ------------------------------------------------------------------------------------------- training legs
def train_leg(handler, device, rank, world, steps, warmup, barrier, batch, hw, flop_per_px, e2e=True):
    """One data-parallel training configuration: device time per step (CUDA events, max over ranks) and, optionally,
    the same step through the reference-facing call `handler.run_train(x_cpu, y_cpu)` with pinned host buffers."""
    import torch.distributed as dist
    from rumpy_b200 import train_native
    if world > 1:
        handler.set_multi_gpu()
    eng = handler.net.native_engine()
    xh = torch.from_numpy(recipe.make_input((batch, 3, hw, hw), seed=80 + rank)).pin_memory()
    yh = torch.from_numpy(recipe.make_input((batch, 3, hw * SCALE, hw * SCALE), seed=180 + rank)).pin_memory()
    xd, yd = xh.to(device), yh.to(device)
    losses = []
    for _ in range(max(warmup, 3)):
        losses.append(train_native.train_step(handler.net, handler.optimizer, xd, yd, allreduce=handler._ddp)[0])
    barrier()
    evs = []
    for _ in range(steps):
        e0, e1 = ev(), ev()
        e0.record()
        losses.append(train_native.train_step(handler.net, handler.optimizer, xd, yd, allreduce=handler._ddp)[0])
        e1.record()
        evs.append((e0, e1))
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    e2e_ms, out_cpu = float('nan'), None
    if e2e:
        for _ in range(3):     # both pinned result buffers exist before timing
            loss_np, out_cpu = handler.run_train(x=xh, y=yh)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            loss_np, out_cpu = handler.run_train(x=xh, y=yh)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    launches = eng.lib.rumpy_net_num_launches(eng.handle) + eng.lib.rumpy_net_num_launches_backward(eng.handle) + 4
    flop = 3 * flop_per_px * batch * hw * hw
    pk = peaks()
    res = {'value': world * batch / (float(t[0]) * 1e-3), 'unit': 'patches/s', 'ms_per_step': float(t[0]),
           'n_gpus': world, 'tflops_per_gpu': flop / (float(t[0]) * 1e-3) * 1e-12,
           'frac_of_sustained_peak': flop / (float(t[0]) * 1e-3) * 1e-12 / pk['tflops_sustained'],
           'gpu_launches_per_step': int(launches), 'loss_first': float(losses[0]), 'loss_last': float(losses[-1]),
           'grad_bytes': int(eng.flat_params.numel() * 4)}
    if e2e:
        res['e2e'] = {'value': world * batch / (float(t[1]) * 1e-3), 'unit': 'patches/s',
                      'h2d_bytes_per_step': int(xh.numel() * 4 + yh.numel() * 4),
                      'd2h_bytes_per_step': int(out_cpu.numel() * 4 + 4),
                      'api': 'handler.run_train(x_cpu, y_cpu) -> (loss numpy, SR batch cpu)'}
    return res


def dp_check(handler, device, rank, world):
    """N > 1: the gradient the product path hands to Adam (chunked NCCL all-reduce overlapped with the backward, times
    1/world) against the mean of the ranks' own gradients, reduced separately in float64."""
    import torch.distributed as dist
    from rumpy_b200 import train_native
    eng = handler.net.native_engine()
    x = torch.from_numpy(recipe.make_input((TRAIN_BATCH, 3, TRAIN_HW, TRAIN_HW), seed=80 + rank)).to(device)
    y = torch.from_numpy(recipe.make_input((TRAIN_BATCH, 3, TRAIN_HW * SCALE, TRAIN_HW * SCALE), seed=180 + rank)).to(device)
    out = eng.forward(x, training=True)
    _, dy = train_native.l1_loss(out, y, want_grad=True)
    eng.backward(x, dy)
    own = eng.flat_grads.clone()
    mean = own.double()
    dist.all_reduce(mean, op=dist.ReduceOp.SUM)
    mean = (mean / world).float()
    out = eng.forward(x, training=True)
    _, dy = train_native.l1_loss(out, y, want_grad=True)
    chunks = eng.backward_chunks()
    eng.backward(x, dy)
    handler._ddp.chunked(eng.flat_grads, chunks)
    torch.cuda.synchronize()
    err = float((eng.flat_grads / world - mean).abs().max() / mean.abs().max())
    differs = float((own - mean).abs().max() / mean.abs().max())     # ranks really hold different batches
    t = torch.tensor([err], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {'rel_err_vs_mean_of_rank_grads': float(t[0]), 'own_vs_mean_rel_diff': differs,
            'ok': bool(float(t[0]) <= 1e-5 and differs > 1e-3)}


def frame_leg(device, rank, world, barrier):
    """BASELINE configs[4]: RCAN x4 on whole 1920x1080 frames (-> 7680x4320), frames sharded round-robin over the ranks,
    no collective.  4 timed frames per rank after a warm-up pair, measured twice: one frame at a time and two frames in
    flight on two streams (parallel.FramesInFlight: one frame's HBM-bound channel-attention passes can co-run with the
    other's convs -- 147.6 vs 159.0 ms with the bf16 pre-attention activation, 174.8 vs 168.7 ms with the fp32 one); value
    = frames of all ranks / max-over-ranks time of the faster schedule, both are reported."""
    import torch.distributed as dist
    from rumpy_b200 import parallel
    from rumpy_b200.SISR.models.advanced.architectures import RCAN
    net = RCAN()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in make_state_dict().items()}, strict=True)
    net = net.to(device).eval()
    per_rank = 4
    mine = parallel.shard_round_robin(range(per_rank * world), rank, world)
    frames = [torch.rand((1, 3, 1080, 1920), device=device, generator=torch.Generator(device).manual_seed(i)) for i in mine]
    res = {}
    for depth in (2, 1):
        pipe = parallel.FramesInFlight(net, depth=depth)
        sink = lambda i, out: None                           # outputs are dropped (a real caller encodes / copies them)
        pipe.run(frames[:2], consume=sink)
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        pipe.run(frames, consume=sink)
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[depth] = float(t[0]) / per_rank
        mode = pipe.engines[0].lib.rumpy_net_trunk_mode(pipe.engines[0].handle)
        del pipe
        torch.cuda.empty_cache()
    del net, frames
    torch.cuda.empty_cache()
    best = 2 if res[2] < res[1] else 1
    ms_frame = res[best]
    tflops = FLOP_PER_LR_PIXEL * 1080 * 1920 / ms_frame * 1e-9
    return {'value': world * 4320 * 7680 / ms_frame * 1e-3, 'unit': 'Mpix/s', 'ms_per_frame_per_gpu': ms_frame,
            'n_gpus': world, 'frames_timed': per_rank * world, 'frames_in_flight_per_gpu': best,
            'tflops_per_gpu': tflops, 'frac_of_sustained_peak': tflops / peaks()['tflops_sustained'],
            'one_at_a_time_ms_per_frame': res[1], 'two_in_flight_ms_per_frame': res[2], 'trunk_mode': int(mode),
            'note': 'whole frames, no tiling (CALayer pools the full image), round-robin over ranks, no collective'}


# ------------------------------------------------------------------------------------------- extras (N = 1, side file)
def extra_configs(device):
    """SURVEY 8(f) widening rows and the like-for-like GPU baselines; informational, written to a side file."""
    from rumpy_b200 import train_native
    from rumpy_b200.optim import FusedAdam
    out = {}

    def timed(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / iters
    from rumpy_b200.SISR.models.attention_manipulators.architectures import QRCAN
    qnet = QRCAN(style='standard', num_metadata=10, include_q_layer=True)
    qspec = [(k, tuple(v.shape)) for k, v in qnet.state_dict().items()]
    qnet.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(qspec, seed=8).items()}, strict=True)
    qnet = qnet.to(device).eval()
    x = torch.rand((BATCH, 3, LR_HW, LR_HW), device=device)
    meta = torch.rand((BATCH, 10, 1, 1), device=device)
    with torch.no_grad():
        ms = timed(lambda: qnet(x, meta), 10)
    out['qrcan_x4_infer'] = {'ms_per_step': ms, 'out_mpix_per_s': OUT_MPIX_PER_STEP / (ms * 1e-3)}
    del qnet
    from rumpy_b200.SISR.models.advanced.architectures import HAN
    hnet = HAN()
    hsd = recipe.make_weights(recipe.han_spec(20), seed=8)
    hsd['la.gamma'] = np.array([0.3], dtype=np.float32)
    hsd['csa.gamma'] = np.array([0.5], dtype=np.float32)
    hnet.load_state_dict({k: torch.from_numpy(v) for k, v in hsd.items()}, strict=True)
    hnet = hnet.to(device).eval()
    with torch.no_grad():
        ms = timed(lambda: hnet(x), 10)
    out['han_x4_infer'] = {'ms_per_step': ms, 'out_mpix_per_s': OUT_MPIX_PER_STEP / (ms * 1e-3)}
    xt = torch.rand((16, 3, 64, 64), device=device)
    yt = torch.rand((16, 3, 256, 256), device=device)
    mt = torch.rand((16, 10, 1, 1), device=device)
    for key, tnet, md in (('han_x4_train', hnet, None),
                          ('qrcan_x4_train', QRCAN(style='standard', num_metadata=10, include_q_layer=True), mt)):
        if md is not None:
            tnet.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(qspec, seed=8).items()}, strict=True)
            tnet = tnet.to(device)
        tnet.train()
        topt = FusedAdam(list(tnet.parameters()), lr=1e-4)
        ms = timed(lambda: train_native.train_step(tnet, topt, xt, yt, metadata=md), 5)
        out[key] = {'ms_per_step': ms, 'patches_per_s_per_gpu': 16 / ms * 1e3}
        del tnet, topt
        torch.cuda.empty_cache()
    del hnet
    torch.cuda.empty_cache()
    out['torch_eager_gpu_baseline'] = eager_gpu_baseline(device)
    out['glue'] = glue_bench(device)
    return out


def glue_bench(device):
    """SURVEY 8(f) ranks 3 and 4 (csrc/glue.cu): eval glue (PSNR(Y), uint8 quantise, bicubic baseline) and the
    training-patch pipeline on the device, each against its HBM roofline."""
    from rumpy_b200.shared_framework.data import (DevicePairSet, PairSet, bicubic_upsample_device, psnr_y_device,
                                                  quantize_u8_device)
    pk = peaks()
    res = {}

    def timed(fn, iters=20):
        for _ in range(3):
            fn()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / iters

    def row(ms, nbytes):
        return {'ms': ms, 'gb_per_s': nbytes / ms * 1e-6, 'frac_of_hbm_peak': nbytes / ms * 1e-6 / pk['hbm'],
                'algorithmic_bytes': nbytes}
    fsr = torch.rand((1, 3, 4320, 7680), device=device)
    fhr = torch.rand((1, 3, 4320, 7680), device=device)
    res['psnr_y_frame_4320x7680'] = row(timed(lambda: psnr_y_device(fsr, fhr), 10), 2 * fsr.numel() * 4)
    res['quantize_u8_frame_4320x7680'] = row(timed(lambda: quantize_u8_device(fsr), 10), fsr.numel() * 5)
    del fsr, fhr
    flr = torch.rand((1, 3, 1080, 1920), device=device)
    res['bicubic_upsample_frame_1080p_x4'] = row(timed(lambda: bicubic_upsample_device(flr, 4), 10), flr.numel() * 4 * 17)
    del flr
    torch.cuda.empty_cache()
    cfg = {'synthetic': 64, 'crop': 64, 'random_augment': True}
    host, dev = PairSet(cfg, SCALE, seed=8), DevicePairSet(cfg, SCALE, seed=8, device=device.index or 0)
    for _ in dev.batches(16):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nb = sum(1 for _ in host.batches(16))
    host_s = time.perf_counter() - t0
    e0, e1 = ev(), ev()
    e0.record()
    nbd = sum(1 for _ in dev.batches(16))
    e1.record()
    e1.synchronize()
    res['patch_pipeline'] = {'host_patches_per_s': nb * 16 / host_s,
                             'device_patches_per_s': nbd * 16 / (e0.elapsed_time(e1) * 1e-3)}
    return res


def eager_gpu_baseline(device):
    """Like-for-like GPU baseline: the reference's arithmetic as PyTorch eager (cuDNN) on the SAME B200 --
    the oracle's functional restatement of RCAN.forward / run_train moved to the device, in fp32, TF32 and
    bf16-autocast.  Baseline leg only (reported beside cpu_baseline, never the measured product)."""
    from oracle import sr_torch_cpu
    sd = {k: torch.from_numpy(v).to(device) for k, v in make_state_dict().items()}
    x = torch.rand((BATCH, 3, LR_HW, LR_HW), device=device)
    res = {}

    def timed(fn, iters):
        for _ in range(2):
            fn()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / iters
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for mode in ('fp32', 'tf32', 'bf16_autocast'):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = (mode == 'tf32')

            def fwd():
                with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16, enabled=(mode == 'bf16_autocast')):
                    return sr_torch_cpu.rcan_forward(sd, x, 10, 20, SCALE)
            ms = timed(fwd, 5)
            res['infer_' + mode] = {'ms_per_step': ms, 'out_mpix_per_s': OUT_MPIX_PER_STEP / (ms * 1e-3)}
        xt = torch.rand((16, 3, 64, 64), device=device)
        yt = torch.rand((16, 3, 256, 256), device=device)
        for mode in ('tf32', 'bf16_autocast'):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = (mode == 'tf32')
            params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            opt = torch.optim.Adam(list(params.values()), lr=1e-4)

            def step():
                with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(mode == 'bf16_autocast')):
                    o = sr_torch_cpu.rcan_forward(params, xt, 10, 20, SCALE)
                loss = torch.nn.functional.l1_loss(o.float(), yt)
                opt.zero_grad()
                loss.backward()
                opt.step()
            ms = timed(step, 3)
            res['train_' + mode] = {'ms_per_step': ms, 'patches_per_s': 16 / (ms * 1e-3)}
            del params, opt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    res['note'] = 'torch %s eager, cuDNN, same weights and shapes as configs[1] / configs[2]' % torch.__version__
    return res


# ------------------------------------------------------------------------------------------- main arm
def run_b200(args, rank, world):
    import torch.distributed as dist
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    from rumpy_b200.shared_framework.models import define_model

    tmp = tempfile.mkdtemp()
    pk = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    handler = define_model('rcan', device=local_rank, model_save_dir=tmp, eval_mode=True, scale=SCALE)
    handler.net.load_state_dict({k: torch.from_numpy(v) for k, v in make_state_dict().items()}, strict=True)
    handler.net.eval()
    eng = handler.net.native_engine()
    x_host = torch.from_numpy(recipe.make_input((BATCH, 3, LR_HW, LR_HW), seed=8 + rank)).pin_memory()
    x_dev = x_host.to(device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    # ---- inference, device-resident steps
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            eng.forward_graphed(x_dev)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    evs = []
    with torch.no_grad():
        for _ in range(args.steps):
            flush.zero_()                                    # L2 flush between timed iterations (untimed)
            e0, e1 = ev(), ev()
            e0.record()
            eng.forward_graphed(x_dev)
            e1.record()
            evs.append((e0, e1))
    barrier()
    infer_total_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    # ---- inference end to end through the reference-facing handler call, HOST buffers in and out
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out_cpu, _, _ = handler.run_eval(x_host)         # both pinned output buffers exist before timing starts
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_cpu, _, _ = handler.run_eval(x_host)
        torch.cuda.synchronize()
        infer_e2e_s = time.perf_counter() - t0
    infer_launches = int(eng.lib.rumpy_net_num_launches(eng.handle))
    roof = trunk_kernel_roofline(eng, x_dev, flush, pk)
    roof_conv = conv_kernel_roofline(device, pk) if rank == 0 else None
    if roof is None and roof_conv is not None:
        roof = {'bound': 'tensor', 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'traffic': None, **roof_conv}
    del handler, eng
    torch.cuda.empty_cache()

    # ---- training legs (data parallel at world > 1), same run, same clocks record
    tr = full = frame = check = None
    tsteps = max(3, min(args.steps, 10))
    if not args.no_train:
        h = define_model('rcan', device=local_rank, model_save_dir=tmp, eval_mode=False, lr=1e-4, scale=SCALE)
        h.net.load_state_dict({k: torch.from_numpy(v) for k, v in make_state_dict().items()}, strict=True)
        tr = train_leg(h, device, rank, world, args.steps if args.metric == 'train' else tsteps, args.warmup, barrier,
                       TRAIN_BATCH, TRAIN_HW, FLOP_PER_LR_PIXEL)
        if world > 1:
            check = dp_check(h, device, rank, world)
        del h
        torch.cuda.empty_cache()
    if not args.no_configs:
        # configs[3]: EDSR x4 full (32 ResBlocks, 256 ch, res_scale 0.1), bf16 operands, batch 16 x 64x64 per GPU
        # (BASELINE.json gives no batch / patch size: configs[2]'s are used), data parallel -- 172 MB of gradient
        h = define_model('edsr', device=local_rank, model_save_dir=tmp, eval_mode=False, lr=1e-4, scale=SCALE,
                         num_features=256, num_blocks=32, res_scale=0.1)
        spec = recipe.edsr_spec(32, 256, 4)
        h.net.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(spec, seed=8).items()}, strict=True)
        full = train_leg(h, device, rank, world, 5, 3, barrier, TRAIN_BATCH, TRAIN_HW, EDSR_FULL_FLOP_PER_LR_PIXEL, e2e=False)
        del h
        torch.cuda.empty_cache()
        frame = frame_leg(device, rank, world, barrier)
    clocks = sampler.result()
    extra = None
    if world == 1 and not args.no_extra:
        extra = extra_configs(device)

    t = torch.tensor([infer_total_ms, infer_e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    ms_per_step = float(t[0]) / args.steps
    infer_value = world * OUT_MPIX_PER_STEP / (ms_per_step * 1e-3)
    infer_e2e = {'value': world * OUT_MPIX_PER_STEP / (float(t[1]) / args.steps), 'unit': 'Mpix/s',
                 'h2d_bytes_per_step': int(x_host.numel() * 4), 'd2h_bytes_per_step': int(out_cpu.numel() * 4),
                 'api': 'RCANHandler.run_eval(x_cpu) -> out_cpu'}
    cpu = None
    if world == 1:
        cores = os.cpu_count() or 1
        kind, fwd, _ = reference_cpu_handler(cores)
        cpu_s, cpu_iters = time_cpu(fwd, BATCH, min_seconds=10.0, max_iters=30)
        cpu = {'value': OUT_MPIX_PER_STEP / cpu_s, 'unit': 'Mpix/s', 'cores': cores, 'kind': kind,
               'sample': f'full batch of {BATCH} patches, median of {cpu_iters} forwards, fp32 torch-CPU'
                         + (' through the unmodified reference handler (baseline/_ref)' if kind == 'reference' else '')}
    whole_tflops = FLOP_PER_LR_PIXEL * BATCH * LR_HW * LR_HW / (ms_per_step * 1e-3) * 1e-12
    common = {'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic'}
    cfg_common = {'scale': SCALE, 'weights': 'random init (numpy recipe seed 8)',
                  'compute': 'bf16 operands, fp32 accumulate, fp32 residual stream'}
    if args.metric == 'train' and tr is not None:
        line = {'metric': 'RCAN x4 train patches/s', 'value': tr['value'], 'unit': 'patches/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': tr['ms_per_step'], **common,
                'config': {'workload': TRAIN_WORKLOAD, 'batch_per_gpu': TRAIN_BATCH, 'lr_patch': TRAIN_HW, **cfg_common,
                           'l2': 'every step streams 11 GB of saved activations: working set far above the 126 MB L2',
                           'parallelism': f'dp{world}: NCCL all-reduce of the flat fp32 gradient in 4 chunks overlapped '
                                          'with the weight-gradient kernels'},
                'clocks': clocks, 'e2e': tr['e2e'], 'gpu_launches': tr['gpu_launches_per_step'] * args.steps,
                'launches_per_step': tr['gpu_launches_per_step'],
                'roofline': {'bound': 'tensor', 'kernel': 'whole train step (trunk_pipe + trunk_bwd + wgrad_tc + Adam)',
                             'achieved': tr['tflops_per_gpu'], 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
                             'frac': tr['frac_of_sustained_peak'], 'traffic': None,
                             'peak_source': pk['source'] + ', sustained figure (kernels timed inside a long step)'},
                'loss_first': tr['loss_first'], 'loss_last': tr['loss_last'],
                'infer': {'value': infer_value, 'unit': 'Mpix/s', 'ms_per_step': ms_per_step}}
    else:
        line = {'metric': 'RCAN x4 output Mpix/s (infer)', 'value': infer_value, 'unit': 'Mpix/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, **common,
                'config': {'workload': WORKLOAD, 'batch_per_gpu': BATCH, 'lr_patch': LR_HW, **cfg_common,
                           'l2': 'flushed between timed steps (256 MB memset, untimed)',
                           'parallelism': f'independent patch batches per GPU x{world}, no collective'},
                'clocks': clocks, 'e2e': infer_e2e, 'gpu_launches': infer_launches * args.steps,
                'launches_per_step': infer_launches, 'whole_step_tflops': whole_tflops,
                'whole_step_frac_of_sustained_peak': whole_tflops / pk['tflops_sustained'], 'roofline': roof}
        if tr is not None:
            line['train'] = {'metric': 'RCAN x4 train patches/s', **tr, 'scaling': 'weak',
                             'config': {'workload': TRAIN_WORKLOAD, 'lr': 1e-4}}
    if cpu is not None:
        line['cpu_baseline'] = cpu
    if roof_conv is not None and args.metric != 'train':
        line['roofline_per_layer_conv'] = roof_conv
    if check is not None:
        line['dp_check'] = check
    if full is not None:
        line['edsr_full_train'] = {'metric': 'EDSR x4 full (32 x 256, res_scale 0.1) train patches/s', **full,
                                   'config': 'batch 16 x 64x64 LR patches per GPU (BASELINE gives none), bf16 operands, dp'}
    if frame is not None:
        line['frame_1080p'] = {'metric': 'RCAN x4 1920x1080 -> 7680x4320 output Mpix/s', **frame}
    if extra is not None:
        log('extra_configs: ' + json.dumps(extra))
        try:
            os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
            with open(os.path.join(ROOT, 'gpurun_out', 'bench_extra.json'), 'w') as f:
                json.dump(extra, f, indent=1)
        except OSError:
            pass
    emit(line)


def main():
    # NCCL / torch print banners on stdout; the contract is ONE JSON line there -> park fd 1 on stderr and keep
    # the real stdout for the final line only
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--metric', default='infer', choices=['infer', 'train'],
                    help='which leg is the headline of the JSON line (train = the data-parallel step with the all-reduce)')
    ap.add_argument('--no-train', action='store_true', help='skip the training (configs[2]) leg')
    ap.add_argument('--no-configs', action='store_true', help='skip configs[3]/[4] (EDSR-full DP training, 1080p frames)')
    ap.add_argument('--no-extra', action='store_true', help='skip the informational side-file legs (N = 1 only)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback '
                         '(use --impl reference for the CPU baseline)')
    run_b200(args, rank, world)


if __name__ == '__main__':
    main()
