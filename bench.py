"""Benchmark of the hot path: RCAN x4 (10 groups x 20 RCAB, 64 ch) forward on synthetic 48x48 LR patches,
batch 16 per GPU (BASELINE.json configs[1]); metric = output Mpix/s.

    python bench.py --gpus 1 --steps 50 --warmup 5                 # this repo's sm_100a path
    python bench.py --impl reference --steps 5 --warmup 1          # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N ...                                      # weak scaling: 16 patches per GPU, no collective

One JSON line on stdout (rank 0).  `value` = whole-job throughput with inputs resident in HBM (CUDA-event
device time per step, L2 flushed between steps, max over ranks); `e2e` = the same metric through the
reference-facing handler call `RCANHandler.run_eval` with pinned HOST buffers (H2D + D2H inside the timed
region); `roofline` = the dominant kernel (tcgen05 conv 64->64) timed alone with CUDA events against the
measured bf16 peak; `cpu_baseline` = the oracle port (torch-CPU restatement of the reference) on the host cores.
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
BATCH, LR_HW, SCALE, CH = 16, 48, 4, 64
FLOP_PER_LR_PIXEL = 31835520          # 2*MAC over all convs of RCAN x4
CONV64_FLOP_PER_PIXEL = 2 * 64 * 64 * 9
OUT_MPIX_PER_STEP = BATCH * (LR_HW * SCALE) ** 2 / 1e6


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=p['bf16_tflops'], tflops_sustained=p['bf16_tflops_sustained'], hbm=p['hbm_gbs'],
                    source='MEASURED_PEAKS.json (measured)')
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0, source='B200_PROFILING.md fallback')


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


def cpu_forward_fn(threads):
    """Oracle port = torch-CPU functional restatement of the reference's RCAN.forward (oracle/sr_torch_cpu.py)."""
    from oracle import sr_torch_cpu
    torch.set_num_threads(threads)
    sd = {k: torch.from_numpy(v) for k, v in make_state_dict().items()}

    def fwd(x):
        with torch.no_grad():
            return sr_torch_cpu.rcan_forward(sd, x, 10, 20, SCALE)
    return fwd


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
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port), all host threads."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    fwd = cpu_forward_fn(cores)
    x16 = torch.from_numpy(recipe.make_input((BATCH, 3, LR_HW, LR_HW), seed=8))
    t0 = time.perf_counter()
    fwd(x16)
    est = time.perf_counter() - t0
    # bound the whole run to a few minutes: shrink the per-step sample if one full batch is slow
    budget = 150.0
    batch = BATCH
    while batch > 1 and est * (batch / BATCH) * (args.steps + args.warmup) > budget:
        batch //= 2
    x = x16[:batch].contiguous()
    for _ in range(args.warmup):
        fwd(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fwd(x)
    dt = (time.perf_counter() - t0) / args.steps
    mpix_s = batch * (LR_HW * SCALE) ** 2 / 1e6 / dt
    sample = f'{batch} of {BATCH} patches per step x {args.steps} steps, fp32, torch-CPU (oneDNN)'
    line = {
        'impl': 'reference', 'metric': 'RCAN x4 output Mpix/s (infer)', 'value': mpix_s, 'unit': 'Mpix/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3 * (BATCH / batch),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_gpu': BATCH, 'lr_patch': LR_HW, 'scale': SCALE},
        'cpu_baseline': {'value': mpix_s, 'unit': 'Mpix/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': mpix_s, 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def conv_kernel_roofline(device, pk):
    """Dominant kernel alone: tcgen05 conv 64->64 (+bias+ReLU) at the workload's activation shape, 20 launches in a
    CUDA graph, timed with CUDA events on the launching stream."""
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get('conv64_dram_bytes_per_launch')
    return {'bound': 'tensor', 'kernel': 'conv3x3_tc_kernel<64,resident-B> (64->64, bias+ReLU, 16x48x48)',
            'achieved': achieved, 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'frac': achieved / pk['tflops'],
            'traffic': traffic, 'us_per_launch': best * 1e6, 'flops_per_launch': flops,
            'peak_source': pk['source'] + ', burst figure (kernel timed alone)'}


def trunk_kernel_roofline(eng, x_dev, flush, pk):
    """Dominant kernel of the headline forward: the trunk kernel that runs all 411 body convs (64->64, 3x3) with
    their epilogues.  Timed live with CUDA events recorded by the library right before / after the kernel on the
    launching stream (eager forwards, L2 flushed before each), best of 5."""
    lib = eng.lib
    mode = lib.rumpy_net_trunk_mode(eng.handle)
    if mode <= 0:
        return None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); e1.record()
    torch.cuda.synchronize()
    eng.set_trunk_events(e0, e1)
    best = None
    try:
        with torch.no_grad():
            for _ in range(5):
                flush.zero_()
                eng.forward(x_dev)
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1) * 1e-3
                best = t if best is None else min(best, t)
    finally:
        eng.set_trunk_events(None, None)
    n_convs = 10 * (2 * 20 + 1) + 1
    flops = n_convs * CONV64_FLOP_PER_PIXEL * BATCH * LR_HW * LR_HW
    achieved = flops / best * 1e-12
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get('trunk_dram_bytes_per_launch')
    name = {1: 'trunk_pipe_kernel (persistent dataflow, epoch flags)',
            2: 'trunk_cluster_kernel (one 6-CTA cluster per image, DSMEM halo exchange)',
            3: 'trunk_band_kernel (role-swapped: weights in TMEM, N = 144 pixels per MMA; one 6-CTA cluster of row '
               'bands per image, DSMEM halo rows)'}[mode]
    return {'bound': 'tensor', 'kernel': name + f': {n_convs} fused conv3x3 64->64 layers + CA + skips, 16x48x48',
            'achieved': achieved, 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'frac': achieved / pk['tflops'],
            'traffic': traffic, 'us_per_launch': best * 1e6, 'flops_per_launch': flops,
            'peak_source': pk['source'] + ', burst figure (one ~2.5 ms kernel)'}


def train_bench(handler_factory, device, rank, world, steps, warmup, barrier):
    """BASELINE.json configs[2]: RCAN x4 training fwd/bwd, L1 loss, Adam 1e-4, 64x64 LR patches, batch 16 per GPU,
    data parallel (NCCL gradient all-reduce).  Returns (ms_per_step max-over-ranks, e2e ms, first loss, last loss)."""
    import torch.distributed as dist
    from rumpy_b200 import train_native
    handler = handler_factory()
    if world > 1:
        handler.set_multi_gpu()
    eng = handler.net.native_engine()
    TB, THW = 16, 64
    xh = torch.from_numpy(recipe.make_input((TB, 3, THW, THW), seed=80 + rank)).pin_memory()
    yh = torch.from_numpy(recipe.make_input((TB, 3, THW * SCALE, THW * SCALE), seed=180 + rank)).pin_memory()
    xd, yd = xh.to(device), yh.to(device)
    losses = []
    for _ in range(max(warmup, 3)):
        losses.append(train_native.train_step(handler.net, handler.optimizer, xd, yd, allreduce=handler._ddp)[0])
    barrier()
    evs = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        losses.append(train_native.train_step(handler.net, handler.optimizer, xd, yd, allreduce=handler._ddp)[0])
        e1.record()
        evs.append((e0, e1))
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    # end to end through the reference-facing call: host batch in, loss (numpy) + SR batch on host out
    for _ in range(3):     # keep the result like the timed loop does: both pinned output buffers exist before timing
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
    return dict(dev_ms=float(t[0]), e2e_ms=float(t[1]), loss_first=float(losses[0]), loss_last=float(losses[-1]),
                launches=int(launches), h2d=int(xh.numel() * 4 + yh.numel() * 4), d2h=int(out_cpu.numel() * 4 + 4),
                batch=TB, hw=THW)


def extra_configs(device):
    """BASELINE.json configs[3] and configs[4] on ONE GPU (informational; the headline stays configs[1])."""
    from rumpy_b200 import train_native
    from rumpy_b200.optim import FusedAdam
    from rumpy_b200.SISR.models.advanced.architectures import EDSR, RCAN
    out = {}

    def ev():
        return torch.cuda.Event(enable_timing=True)
    # configs[4]: RCAN x4 whole-frame inference 1920x1080 -> 7680x4320 (frames shard round-robin over GPUs)
    net = RCAN()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in make_state_dict().items()}, strict=True)
    net = net.to(device).eval()
    eng = net.native_engine()
    with torch.no_grad():
        x = torch.rand((1, 3, 1080, 1920), device=device)
        for _ in range(2):
            eng.forward(x)
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(3):
            eng.forward(x)
        e1.record()
        e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out['rcan_x4_frame_1080p'] = {'ms_per_frame': ms, 'out_mpix_per_s': 4320 * 7680 / ms * 1e-3,
                                  'tflops': FLOP_PER_LR_PIXEL * 1080 * 1920 / ms * 1e-9,
                                  'note': 'whole frame, no tiling (CALayer pools the full image)'}
    del net, eng, x
    torch.cuda.empty_cache()
    # configs[3]: EDSR x4 full (32 ResBlocks, 256 ch, res_scale 0.1) training step, 16 x 64x64 LR patches per GPU
    spec = recipe.edsr_spec(32, 256, 4)
    net = EDSR(net_features=256, num_blocks=32, res_scale=0.1)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(spec, seed=8).items()}, strict=True)
    net = net.to(device).train()
    opt = FusedAdam(list(net.parameters()), lr=1e-4)
    x = torch.rand((16, 3, 64, 64), device=device)
    y = torch.rand((16, 3, 256, 256), device=device)
    for _ in range(3):
        train_native.train_step(net, opt, x, y)
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(5):
        train_native.train_step(net, opt, x, y)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out['edsr_full_x4_train'] = {'ms_per_step': ms, 'patches_per_s_per_gpu': 16 / ms * 1e3,
                                 'tflops': 3 * 100505088 * 16 * 64 * 64 / ms * 1e-9,
                                 'note': 'batch 16 x 64x64 LR patches per GPU (BASELINE gives no batch size)'}
    del net, opt
    torch.cuda.empty_cache()
    # Q-RCAN (meta-attention; sample q-rcan.toml: blur-kernel metadata M=10, q-node in every
    # RCAB) at configs[1]'s batch, through QRCAN.forward(x, metadata)
    from rumpy_b200.SISR.models.attention_manipulators.architectures import QRCAN
    qnet = QRCAN(style='standard', num_metadata=10, include_q_layer=True)
    qspec = [(k, tuple(v.shape)) for k, v in qnet.state_dict().items()]
    qnet.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(qspec, seed=8).items()}, strict=True)
    qnet = qnet.to(device).eval()
    x = torch.rand((BATCH, 3, LR_HW, LR_HW), device=device)
    meta = torch.rand((BATCH, 10, 1, 1), device=device)
    with torch.no_grad():
        for _ in range(3):
            qnet(x, meta)
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(10):
            qnet(x, meta)
        e1.record()
        e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out['qrcan_x4_infer'] = {'ms_per_step': ms, 'out_mpix_per_s': OUT_MPIX_PER_STEP / (ms * 1e-3),
                             'launches': int(qnet.native_engine().lib.rumpy_net_num_launches(qnet.native_engine().handle)),
                             'note': f'{BATCH} x {LR_HW}x{LR_HW}, metadata copy + graph replay + output clone per call'}
    del qnet
    torch.cuda.empty_cache()
    # HAN (10 x 20 RCAN groups + layer attention + channel-spatial attention) at configs[1]'s batch
    from rumpy_b200.SISR.models.advanced.architectures import HAN
    hnet = HAN()
    hsd = recipe.make_weights(recipe.han_spec(20), seed=8)
    hsd['la.gamma'] = np.array([0.3], dtype=np.float32)
    hsd['csa.gamma'] = np.array([0.5], dtype=np.float32)
    hnet.load_state_dict({k: torch.from_numpy(v) for k, v in hsd.items()}, strict=True)
    hnet = hnet.to(device).eval()
    x = torch.rand((BATCH, 3, LR_HW, LR_HW), device=device)
    with torch.no_grad():
        for _ in range(3):
            hnet(x)
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(10):
            hnet(x)
        e1.record()
        e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out['han_x4_infer'] = {'ms_per_step': ms, 'out_mpix_per_s': OUT_MPIX_PER_STEP / (ms * 1e-3),
                           'launches': int(hnet.native_engine().lib.rumpy_net_num_launches(hnet.native_engine().handle)),
                           'note': f'{BATCH} x {LR_HW}x{LR_HW}; ops: head, trunk, LAM (3 kernels), last_conv, CSAM+cat, '
                                   'last, 2 upsampler convs, tail'}
    # train steps of the two widened families at configs[2]'s batch (16 x 64x64 LR patches)
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
        for _ in range(3):
            train_native.train_step(tnet, topt, xt, yt, metadata=md)
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(5):
            train_native.train_step(tnet, topt, xt, yt, metadata=md)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[key] = {'ms_per_step': ms, 'patches_per_s_per_gpu': 16 / ms * 1e3}
        del tnet, topt
        torch.cuda.empty_cache()
    del hnet
    torch.cuda.empty_cache()
    out['torch_eager_gpu_baseline'] = eager_gpu_baseline(device)
    out['glue'] = glue_bench(device)
    return out


def glue_bench(device):
    """SURVEY 8(f) ranks 3 and 4 (csrc/glue.cu): eval glue (PSNR(Y), uint8 quantise) and the training-patch pipeline
    on the device, each against its HBM roofline and against the host (numpy) path it replaces."""
    import time as _t
    from rumpy_b200.shared_framework.data import (DevicePairSet, PairSet, bicubic_upsample_device, psnr_y, psnr_y_device,
                                                  quantize_u8_device)
    pk = peaks()
    res = {}

    def ev():
        return torch.cuda.Event(enable_timing=True)

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
    hw = LR_HW * SCALE
    sr = torch.rand((BATCH, 3, hw, hw), device=device) * 1.2 - 0.1
    hr = torch.rand((BATCH, 3, hw, hw), device=device)
    ms = timed(lambda: psnr_y_device(sr, hr))
    nbytes = 2 * sr.numel() * 4
    res['psnr_y'] = {'ms': ms, 'gb_per_s': nbytes / ms * 1e-6, 'frac_of_hbm_peak': nbytes / ms * 1e-6 / pk['hbm'],
                     'algorithmic_bytes': nbytes}
    ms = timed(lambda: quantize_u8_device(sr))
    nbytes = sr.numel() * 5
    res['quantize_u8'] = {'ms': ms, 'gb_per_s': nbytes / ms * 1e-6, 'frac_of_hbm_peak': nbytes / ms * 1e-6 / pk['hbm'],
                          'algorithmic_bytes': nbytes}
    # a whole x4 frame (configs[4]'s output, 4320 x 7680): large enough to be bound by HBM rather than by launches
    fsr = torch.rand((1, 3, 4320, 7680), device=device)
    fhr = torch.rand((1, 3, 4320, 7680), device=device)
    ms = timed(lambda: psnr_y_device(fsr, fhr), 10)
    nbytes = 2 * fsr.numel() * 4
    res['psnr_y_frame_4320x7680'] = {'ms': ms, 'gb_per_s': nbytes / ms * 1e-6,
                                     'frac_of_hbm_peak': nbytes / ms * 1e-6 / pk['hbm'], 'algorithmic_bytes': nbytes}
    ms = timed(lambda: quantize_u8_device(fsr), 10)
    nbytes = fsr.numel() * 5
    res['quantize_u8_frame_4320x7680'] = {'ms': ms, 'gb_per_s': nbytes / ms * 1e-6,
                                          'frac_of_hbm_peak': nbytes / ms * 1e-6 / pk['hbm'], 'algorithmic_bytes': nbytes}
    del fsr, fhr
    # bicubic baseline (Pillow's 8-bit resampler restated on the device): 4 B read per LR + 4 B written per output element
    lrb = torch.rand((BATCH, 3, LR_HW, LR_HW), device=device)
    ms = timed(lambda: bicubic_upsample_device(lrb, SCALE))
    nbytes = lrb.numel() * 4 * (1 + SCALE * SCALE)
    res['bicubic_upsample'] = {'ms': ms, 'gb_per_s': nbytes / ms * 1e-6, 'frac_of_hbm_peak': nbytes / ms * 1e-6 / pk['hbm'],
                               'algorithmic_bytes': nbytes}
    flr = torch.rand((1, 3, 1080, 1920), device=device)
    ms = timed(lambda: bicubic_upsample_device(flr, 4), 10)
    nbytes = flr.numel() * 4 * 17
    res['bicubic_upsample_frame_1080p_x4'] = {'ms': ms, 'gb_per_s': nbytes / ms * 1e-6,
                                              'frac_of_hbm_peak': nbytes / ms * 1e-6 / pk['hbm'],
                                              'algorithmic_bytes': nbytes}
    del flr
    torch.cuda.empty_cache()
    t0 = _t.perf_counter()
    src, hrc = sr.cpu(), hr.cpu()
    for n in range(BATCH):
        psnr_y(src[n:n + 1], hrc[n:n + 1])
    (src.permute(0, 2, 3, 1).numpy() * 255).clip(0, 255).astype(np.uint8)
    res['host_path_ms'] = (_t.perf_counter() - t0) * 1e3     # D2H fp32 + numpy PSNR + numpy quantise (reference flow)
    cfg = {'synthetic': 64, 'crop': 64, 'random_augment': True}
    host, dev = PairSet(cfg, SCALE, seed=8), DevicePairSet(cfg, SCALE, seed=8, device=device.index or 0)
    for _ in dev.batches(16):      # warm-up epoch (allocator, pinned staging)
        pass
    torch.cuda.synchronize()
    t0 = _t.perf_counter()
    nb = sum(1 for _ in host.batches(16))
    host_s = _t.perf_counter() - t0
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    nbd = sum(1 for _ in dev.batches(16))
    e1.record()
    e1.synchronize()
    res['patch_pipeline'] = {'host_patches_per_s': nb * 16 / host_s, 'device_patches_per_s': nbd * 16 / (e0.elapsed_time(e1) * 1e-3),
                             'note': '16 x (64x64 LR + 256x256 HR) patches per batch: crop + flips + transpose + ToTensor; '
                                     'host = numpy on one core, device = one kernel per batch from uint8 images in HBM'}
    return res


def eager_gpu_baseline(device):
    """Like-for-like GPU baseline: the reference's arithmetic as PyTorch eager (cuDNN) on the SAME B200 --
    the oracle's functional restatement of RCAN.forward / run_train moved to the device, in fp32, TF32 and
    bf16-autocast.  Baseline leg only (reported beside cpu_baseline, never the measured product)."""
    from oracle import sr_torch_cpu

    def ev():
        return torch.cuda.Event(enable_timing=True)
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
        for mode in ('fp32', 'tf32', 'bf16_autocast'):
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
    res['note'] = ('torch %s eager, cuDNN, same weights and shapes as configs[1] / configs[2]; bf16_autocast fails the '
                   '1e-2 output tolerance (BASELINE.md bf16 risk probe) and is listed for speed only' % torch.__version__)
    return res


def run_b200(args, rank, world):
    import torch.distributed as dist
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    from rumpy_b200.shared_framework.models import define_model

    tmp = tempfile.mkdtemp()
    handler = define_model('rcan', device=local_rank, model_save_dir=tmp, eval_mode=True, scale=SCALE)
    handler.net.load_state_dict({k: torch.from_numpy(v) for k, v in make_state_dict().items()}, strict=True)
    handler.net.eval()
    eng = handler.net.native_engine()

    x_host = torch.from_numpy(recipe.make_input((BATCH, 3, LR_HW, LR_HW), seed=8 + rank)).pin_memory()
    x_dev = x_host.to(device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps (value)
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
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.forward_graphed(x_dev)
            e1.record()
            evs.append((e0, e1))
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    # ---- end-to-end through the reference-facing handler call, HOST buffers in and out
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            # keep the result like the timed loop does: the pinned-memory allocator then owns both output buffers
            # before timing starts (a cudaHostAlloc of 7 MB inside the loop cost 5 ms once)
            out_cpu, _, _ = handler.run_eval(x_host)
        barrier()
        t0 = time.perf_counter()
        call_ms = []
        for _ in range(args.steps):
            tc = time.perf_counter()
            out_cpu, _, _ = handler.run_eval(x_host)
            call_ms.append((time.perf_counter() - tc) * 1e3)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        print('e2e per-call ms: min %.3f median %.3f max %.3f (call %d of %d)' %
              (min(call_ms), float(np.median(call_ms)), max(call_ms), int(np.argmax(call_ms)), len(call_ms)), file=sys.stderr)
    launches = eng.lib.rumpy_net_num_launches(eng.handle)
    # ---- training (configs[2]) in the same run, same clocks record
    def handler_factory():
        h = define_model('rcan', device=local_rank, model_save_dir=tmp, eval_mode=False, lr=1e-4, scale=SCALE)
        h.net.load_state_dict({k: torch.from_numpy(v) for k, v in make_state_dict().items()}, strict=True)
        return h
    tr = None
    if not args.no_train:
        del handler
        torch.cuda.empty_cache()
        tr = train_bench(handler_factory, device, rank, world, max(3, min(args.steps, 10)), args.warmup, barrier)
    extra = None
    if world == 1 and not args.no_extra:
        extra = extra_configs(device)
    clocks = sampler.result()

    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = total_ms / args.steps
    value = world * OUT_MPIX_PER_STEP / (ms_per_step * 1e-3)
    e2e_value = world * OUT_MPIX_PER_STEP / (e2e_s / args.steps)
    pk = peaks()
    roof_conv = conv_kernel_roofline(device, pk)
    roof = trunk_kernel_roofline(eng, x_dev, flush, pk) if eng is not None else None
    if roof is None:
        roof, roof_conv = roof_conv, None
    trunk_tflops = FLOP_PER_LR_PIXEL * BATCH * LR_HW * LR_HW / (ms_per_step * 1e-3) * 1e-12
    cores = os.cpu_count() or 1
    fwd = cpu_forward_fn(cores)
    cpu_s, cpu_iters = time_cpu(fwd, BATCH, min_seconds=10.0, max_iters=30)
    line = {
        'metric': 'RCAN x4 output Mpix/s (infer)', 'value': value, 'unit': 'Mpix/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_gpu': BATCH, 'lr_patch': LR_HW, 'scale': SCALE,
                   'parallelism': f'independent patch batches per GPU x{world}, no collective',
                   'l2': 'flushed between timed steps (256 MB memset, untimed)',
                   'weights': 'random init (numpy recipe seed 8)', 'compute': 'bf16 operands, fp32 accumulate, '
                   'fp32 residual stream'},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'Mpix/s', 'h2d_bytes_per_step': int(x_host.numel() * 4),
                'd2h_bytes_per_step': int(out_cpu.numel() * 4), 'api': 'RCANHandler.run_eval(x_cpu) -> out_cpu'},
        'gpu_launches': int(launches) * args.steps,
        'launches_per_step': int(launches),
        'whole_step_tflops': trunk_tflops, 'whole_step_frac_of_sustained_peak': trunk_tflops / pk['tflops_sustained'],
        'roofline': roof,
        'cpu_baseline': {'value': OUT_MPIX_PER_STEP / cpu_s, 'unit': 'Mpix/s', 'cores': cores, 'kind': 'port',
                         'sample': f'full batch of {BATCH} patches, median of {cpu_iters} forwards, fp32 torch-CPU'},
    }
    if roof_conv is not None:
        line['roofline_per_layer_conv'] = roof_conv    # the stand-alone conv kernel (used for shapes the trunk kernels skip)
    if extra is not None:
        line['extra_configs'] = extra
    if tr is not None:
        train_flop = 3 * FLOP_PER_LR_PIXEL * tr['batch'] * tr['hw'] * tr['hw']
        line['train'] = {
            'metric': 'RCAN x4 train patches/s', 'value': world * tr['batch'] / (tr['dev_ms'] * 1e-3),
            'unit': 'patches/s', 'ms_per_step': tr['dev_ms'], 'n_gpus': world, 'scaling': 'weak',
            'config': {'workload': 'RCAN x4 train step (fwd + L1 + bwd + Adam), 64x64 LR patches, batch 16 per GPU, '
                                   'data parallel with bucketed NCCL gradient all-reduce', 'lr': 1e-4},
            'tflops_per_gpu': train_flop / (tr['dev_ms'] * 1e-3) * 1e-12,
            'frac_of_sustained_peak': train_flop / (tr['dev_ms'] * 1e-3) * 1e-12 / pk['tflops_sustained'],
            'e2e': {'value': world * tr['batch'] / (tr['e2e_ms'] * 1e-3), 'unit': 'patches/s',
                    'h2d_bytes_per_step': tr['h2d'], 'd2h_bytes_per_step': tr['d2h'],
                    'api': 'RCANHandler.run_train(x_cpu, y_cpu) -> (loss numpy, SR batch cpu)'},
            'gpu_launches_per_step': tr['launches'], 'loss_first': tr['loss_first'], 'loss_last': tr['loss_last'],
        }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument('--no-train', action='store_true', help='skip the training (configs[2]) section')
    ap.add_argument('--no-extra', action='store_true', help='skip configs[3]/[4] (EDSR-full train, 1080p frame)')
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
