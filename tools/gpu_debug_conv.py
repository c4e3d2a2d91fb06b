"""Bring-up diagnostics for the tcgen05 conv kernel.  Run on the GPU box:

    python tools/gpu_debug_conv.py            # every case, each in its own subprocess (a trap cannot cascade)
    python tools/gpu_debug_conv.py <case>     # one case in-process

Reference = torch fp32 conv2d on the same bf16-rounded operands (TF32 off).  Writes gpurun_out/debug_conv.log.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _imports():
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from rumpy_b200 import ops
    return torch, F, ops


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def report(name, got, ref, tol):
    import torch
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    mx = err.max().item()
    rel = mx / max(ref.abs().max().item(), 1e-12)
    nbad = int((err > tol).sum().item())
    ok = mx <= tol and bool(torch.isfinite(got).all())
    print(f'[{"OK" if ok else "FAIL"}] {name}: max_abs_err={mx:.3e} rel={rel:.3e} ref_absmax={ref.abs().max().item():.3e} '
          f'bad={nbad}/{err.numel()} tol={tol:g}', flush=True)
    if not ok:
        idx = torch.nonzero(err > tol)[:8]
        for i in idx:
            i = tuple(i.tolist())
            print(f'      at {i}: got {got[i].item():.6f} ref {ref[i].item():.6f}')
    return ok


def _conv_case(N, H, W, Cin, Cout, *, wkind='rand', relu=False, bias=True, residual=False, mask=False,
               out_f32=False, out_bf16=True, pool=False, alpha=1.0, seed=0):
    torch, F, ops = _imports()
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = (torch.rand((N, Cin, H, W), generator=g, device='cuda') * 2 - 1)
    xb = nhwc(x).to(torch.bfloat16)
    if wkind == 'rand':
        w = (torch.rand((Cout, Cin, 3, 3), generator=g, device='cuda') * 2 - 1) / (Cin * 9) ** 0.5
    else:  # identity on tap (ky,kx)
        ky, kx = wkind
        w = torch.zeros((Cout, Cin, 3, 3), device='cuda')
        for o in range(min(Cout, Cin)):
            w[o, o, ky, kx] = 1.0
    b = (torch.rand((Cout,), generator=g, device='cuda') - 0.5) if bias else None
    res = torch.rand((N, H, W, Cout), generator=g, device='cuda') if residual else None
    msk = (torch.rand((N, H, W, Cout), generator=g, device='cuda') - 0.5).to(torch.bfloat16) if mask else None
    wp = ops.pack_conv3x3(w)
    yb = torch.full((N, H, W, Cout), float('nan'), dtype=torch.bfloat16, device='cuda') if out_bf16 else None
    yf = torch.full((N, H, W, Cout), float('nan'), dtype=torch.float32, device='cuda') if out_f32 else None
    pp = torch.full((N, ops.pool_rows(H, W, Cout), Cout), float('nan'), device='cuda') if pool else None
    ops.conv3x3(xb, wp, b, residual=res, mask=msk, out_bf16=yb, out_f32=yf, pool_partial=pp, N=N, H=H, W=W,
                Cin=Cin, Cout=Cout, relu=relu, alpha=alpha)
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(xb.float()), w.to(torch.bfloat16).float(), b, padding=1)
    if relu:
        ref = ref.relu()
    ref = ref * alpha
    ref = nhwc(ref)
    if mask:
        ref = ref * (msk.float() > 0)
    if residual:
        ref = ref + res
    ok = True
    scale = max(ref.abs().max().item(), 1.0)
    if out_f32:
        ok &= report('f32 out', yf, ref, 2e-4 * scale)
    if out_bf16:
        ok &= report('bf16 out', yb, ref, 1e-2 * scale)
    if pool:
        src = yf if out_f32 else yb.float()
        got = pp.view(N, -1, Cout).sum(1)
        ok &= report('pool', got, src.sum(dim=(1, 2)), 1e-3 * H * W)
    return ok


CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


@case
def c01_identity_center():
    return _conv_case(1, 8, 16, 64, 64, wkind=(1, 1), bias=False)


@case
def c02_identity_tap00():
    return _conv_case(1, 8, 16, 64, 64, wkind=(0, 0), bias=False)


@case
def c03_identity_tap22_multi_tile():
    return _conv_case(2, 16, 32, 64, 64, wkind=(2, 2), bias=False)


@case
def c04_rand_single_tile():
    return _conv_case(1, 8, 16, 64, 64)


@case
def c05_rand_ragged_relu():
    return _conv_case(2, 13, 21, 64, 64, relu=True)


@case
def c06_f32_residual_alpha():
    return _conv_case(2, 13, 21, 64, 64, residual=True, out_f32=True, alpha=0.1)


@case
def c07_mask():
    return _conv_case(2, 9, 17, 64, 64, mask=True, bias=False)


@case
def c08_pool_f32():
    return _conv_case(2, 13, 21, 64, 64, out_f32=True, out_bf16=False, pool=True)


@case
def c09_pool_bf16():
    return _conv_case(2, 13, 21, 64, 64, pool=True)


@case
def c10_cout256():
    return _conv_case(1, 11, 19, 64, 256)


@case
def c11_cin256_cout256():
    return _conv_case(1, 11, 19, 256, 256, relu=True)


@case
def c12_cin256_cout64():
    return _conv_case(1, 11, 19, 256, 64)


@case
def c13_many_tiles():
    return _conv_case(16, 48, 48, 64, 64, relu=True)


@case
def c20_shuffle_store():
    torch, F, ops = _imports()
    ok = True
    for r, C in ((2, 64), (3, 64), (2, 256)):
        N, H, W = 2, 7, 10
        g = torch.Generator(device='cuda').manual_seed(r)
        x = torch.rand((N, C, H, W), generator=g, device='cuda') * 2 - 1
        xb = nhwc(x).to(torch.bfloat16)
        cout = C * r * r
        w = (torch.rand((cout, C, 3, 3), generator=g, device='cuda') * 2 - 1) / (C * 9) ** 0.5
        b = torch.rand((cout,), generator=g, device='cuda') - 0.5
        wp = ops.pack_conv3x3(w, shuffle_r=r)
        bp = ops.pack_bias(b, shuffle_r=r)
        y = torch.full((N, H * r, W * r, C), float('nan'), dtype=torch.bfloat16, device='cuda')
        ops.conv3x3(xb, wp, bp, out_bf16=y, N=N, H=H, W=W, Cin=C, Cout=cout, out_shuffle_r=r)
        torch.cuda.synchronize()
        ref = F.pixel_shuffle(F.conv2d(nchw(xb.float()), w.to(torch.bfloat16).float(), b, padding=1), r)
        ok &= report(f'shuffle r={r} C={C}', y, nhwc(ref), 1e-2 * max(1.0, ref.abs().max().item()))
    return ok


@case
def c21_unshuffle_load_dgrad():
    """dX of conv+PixelShuffle: g [N,2H,2W,C] -> unshuffle -> conv_transpose with W."""
    torch, F, ops = _imports()
    ok = True
    for r, C in ((2, 64), (3, 64)):
        N, H, W = 1, 9, 12
        gen = torch.Generator(device='cuda').manual_seed(10 + r)
        cout = C * r * r
        w = (torch.rand((cout, C, 3, 3), generator=gen, device='cuda') * 2 - 1) / (C * 9) ** 0.5
        gout = torch.rand((N, C, H * r, W * r), generator=gen, device='cuda') * 2 - 1
        gb = nhwc(gout).to(torch.bfloat16)
        wd = ops.pack_conv3x3(w, shuffle_r=r, dgrad=True)
        dx = torch.full((N, H, W, C), float('nan'), dtype=torch.float32, device='cuda')
        ops.conv3x3(gb, wd, None, out_f32=dx, N=N, H=H, W=W, Cin=cout, Cout=C, in_unshuffle_r=r)
        torch.cuda.synchronize()
        gz = F.pixel_unshuffle(nchw(gb.float()), r)
        ref = F.conv_transpose2d(gz, w.to(torch.bfloat16).float(), padding=1)
        ok &= report(f'unshuffle-dgrad r={r}', dx, nhwc(ref), 2e-4 * max(1.0, ref.abs().max().item()))
    return ok


@case
def c22_dgrad_plain():
    torch, F, ops = _imports()
    N, H, W, C = 2, 10, 18, 64
    gen = torch.Generator(device='cuda').manual_seed(5)
    w = (torch.rand((C, C, 3, 3), generator=gen, device='cuda') * 2 - 1) / (C * 9) ** 0.5
    g = torch.rand((N, C, H, W), generator=gen, device='cuda') * 2 - 1
    gb = nhwc(g).to(torch.bfloat16)
    wd = ops.pack_conv3x3(w, dgrad=True)
    dx = torch.full((N, H, W, C), float('nan'), dtype=torch.float32, device='cuda')
    ops.conv3x3(gb, wd, None, out_f32=dx, N=N, H=H, W=W, Cin=C, Cout=C)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(nchw(gb.float()), w.to(torch.bfloat16).float(), padding=1)
    return report('dgrad', dx, nhwc(ref), 2e-4 * max(1.0, ref.abs().max().item()))


@case
def c30_tail_thin():
    torch, F, ops = _imports()
    ok = True
    for C in (64, 256):
        N, H, W = 2, 21, 37
        gen = torch.Generator(device='cuda').manual_seed(C)
        x = torch.rand((N, C, H, W), generator=gen, device='cuda') * 2 - 1
        xb = nhwc(x).to(torch.bfloat16)
        w = (torch.rand((3, C, 3, 3), generator=gen, device='cuda') * 2 - 1) / (C * 9) ** 0.5
        b = torch.rand((3,), generator=gen, device='cuda') - 0.5
        wp = ops.pack_conv3x3(w, rows_padded=16)
        bp = ops.pack_bias(b, rows_padded=16)
        y = torch.full((N, 3, H, W), float('nan'), device='cuda')
        ops.conv3x3_tail(xb, wp, bp, y, N=N, H=H, W=W, Cin=C, cout_real=3)
        torch.cuda.synchronize()
        ref = F.conv2d(nchw(xb.float()), w.to(torch.bfloat16).float(), b, padding=1)
        ok &= report(f'tail C={C}', y, ref, 2e-4 * max(1.0, ref.abs().max().item()))
    return ok


@case
def c31_head():
    torch, F, ops = _imports()
    ok = True
    for C in (64, 256):
        N, H, W = 2, 13, 21
        gen = torch.Generator(device='cuda').manual_seed(C + 1)
        x = torch.rand((N, 3, H, W), generator=gen, device='cuda')
        w = (torch.rand((C, 3, 3, 3), generator=gen, device='cuda') * 2 - 1) / 27 ** 0.5
        b = torch.rand((C,), generator=gen, device='cuda') - 0.5
        yf = torch.full((N, H, W, C), float('nan'), device='cuda')
        yb = torch.full((N, H, W, C), float('nan'), dtype=torch.bfloat16, device='cuda')
        ops.head_conv(x, w, b, yf, yb)
        torch.cuda.synchronize()
        ref = nhwc(F.conv2d(x, w, b, padding=1))
        ok &= report(f'head f32 C={C}', yf, ref, 1e-5)
        ok &= report(f'head bf16 C={C}', yb, ref, 1e-2)
    return ok


@case
def c32_ca_apply():
    torch, F, ops = _imports()
    ok = True
    for u_f32 in (True, False):
        N, H, W, C, Cr = 3, 13, 21, 64, 4
        gen = torch.Generator(device='cuda').manual_seed(9)
        xin = torch.rand((N, H, W, C), generator=gen, device='cuda')
        u32 = torch.rand((N, H, W, C), generator=gen, device='cuda') * 2 - 1
        u = u32 if u_f32 else u32.to(torch.bfloat16)
        w1 = torch.rand((Cr, C), generator=gen, device='cuda') - 0.5
        b1 = torch.rand((Cr,), generator=gen, device='cuda') - 0.5
        w2 = torch.rand((C, Cr), generator=gen, device='cuda') - 0.5
        b2 = torch.rand((C,), generator=gen, device='cuda') - 0.5
        # pool partials as the conv epilogue would emit them: here simply spread the sums over the partial slots
        tiles = ops.pool_rows(H, W, C) // 2
        pp = torch.zeros((N * tiles, 2, C), device='cuda')
        pp.view(N, tiles * 2, C)[:, 0, :] = u.float().sum(dim=(1, 2))
        xo = torch.empty_like(xin)
        xob = torch.empty((N, H, W, C), dtype=torch.bfloat16, device='cuda')
        sm, sh, sy = torch.empty((N, C), device='cuda'), torch.empty((N, Cr), device='cuda'), torch.empty((N, C), device='cuda')
        ops.ca_apply(pp, u, xin, w1, b1, w2, b2, xo, xob, N=N, H=H, W=W, C=C, save=(sm, sh, sy))
        torch.cuda.synchronize()
        mean = u.float().mean(dim=(1, 2))
        hid = (mean @ w1.T + b1).relu()
        y = torch.sigmoid(hid @ w2.T + b2)
        ref = xin + u.float() * y[:, None, None, :]
        ok &= report(f'ca_apply f32={u_f32}', xo, ref, 1e-5)
        ok &= report('ca_apply bf16', xob, ref, 1e-2)
        ok &= report('ca y', sy, y, 1e-5)
        ok &= report('ca hid', sh, hid, 1e-5)
    return ok


def _wgrad_case(N, H, W, Cin, Cout, r=1, alpha=1.0, accumulate=False, seed=0):
    torch, F, ops = _imports()
    gen = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.rand((N, Cin, H, W), generator=gen, device='cuda') * 2 - 1
    xb = nhwc(x).to(torch.bfloat16)
    gout = torch.rand((N, Cout // (r * r), H * r, W * r), generator=gen, device='cuda') * 2 - 1
    gb = nhwc(gout).to(torch.bfloat16)
    dw = torch.full((Cout, Cin, 3, 3), 0.5 if accumulate else float('nan'), device='cuda')
    ops.conv3x3_wgrad(gb, xb, dw, N=N, H=H, W=W, Cin=Cin, Cout=Cout, g_unshuffle_r=r, alpha=alpha,
                      accumulate=accumulate)
    torch.cuda.synchronize()
    w = torch.zeros((Cout, Cin, 3, 3), device='cuda', requires_grad=True)
    xin = nchw(xb.float())
    y = F.conv2d(xin, w, None, padding=1)
    if r > 1:
        y = F.pixel_shuffle(y, r)
    y.backward(nchw(gb.float()))
    ref = w.grad * alpha + (0.5 if accumulate else 0.0)
    return report(f'wgrad N{N} {H}x{W} {Cin}->{Cout} r={r}', dw, ref, 2e-4 * max(1.0, ref.abs().max().item()))


@case
def w01_wgrad_single_tile():
    return _wgrad_case(1, 8, 16, 64, 64)


@case
def w02_wgrad_ragged():
    return _wgrad_case(2, 13, 21, 64, 64, alpha=0.5, accumulate=True)


@case
def w03_wgrad_many_tiles():
    return _wgrad_case(16, 48, 48, 64, 64)


@case
def w04_wgrad_256():
    return _wgrad_case(1, 11, 19, 256, 128)


@case
def w05_wgrad_unshuffle():
    ok = _wgrad_case(2, 7, 10, 64, 256, r=2)
    ok &= _wgrad_case(1, 5, 9, 64, 576, r=3)
    return ok


@case
def p02_perf_wgrad():
    torch, F, ops = _imports()
    for (N, H, W) in ((16, 48, 48), (16, 64, 64)):
        C = 64
        x = torch.rand((N, H, W, C), device='cuda').to(torch.bfloat16)
        g = torch.rand((N, H, W, C), device='cuda').to(torch.bfloat16)
        dw = torch.empty((C, C, 3, 3), device='cuda')
        for _ in range(3):
            ops.conv3x3_wgrad(g, x, dw, N=N, H=H, W=W, Cin=C, Cout=C)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv3x3_wgrad(g, x, dw, N=N, H=H, W=W, Cin=C, Cout=C)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(f'[PERF] wgrad64 N{N} {H}x{W}: {us:.2f} us/call (op-level, incl. host job upload + sync)', flush=True)
    return True


@case
def p01_perf_conv64():
    """Quick device-time numbers (CUDA events), conv 64->64 + ReLU, config #2 / #3 / big shapes."""
    torch, F, ops = _imports()
    for (N, H, W) in ((16, 48, 48), (16, 64, 64), (1, 1080, 1920)):
        C = 64
        x = torch.rand((N, H, W, C), device='cuda').to(torch.bfloat16)
        w = torch.rand((C, C, 3, 3), device='cuda') - 0.5
        b = torch.rand((C,), device='cuda')
        wp = ops.pack_conv3x3(w)
        y = torch.empty_like(x)
        for _ in range(5):
            ops.conv3x3(x, wp, b, out_bf16=y, N=N, H=H, W=W, Cin=C, Cout=C, relu=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 50
        e0.record()
        for _ in range(iters):
            ops.conv3x3(x, wp, b, out_bf16=y, N=N, H=H, W=W, Cin=C, Cout=C, relu=True)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        flops = 2.0 * N * H * W * C * C * 9
        print(f'[PERF] conv64 N{N} {H}x{W}: {us:.2f} us/launch (incl. host tensor-map encode)  '
              f'{flops / us * 1e-6:.1f} TFLOP/s', flush=True)
    return True


def worker(names):
    """Runs cases in-process; a CUDA trap kills this worker and the parent restarts after the culprit."""
    for name in names:
        print(f'BEGIN {name}', flush=True)
        t0 = time.time()
        try:
            ok = CASES[name]()
        except Exception as e:  # noqa
            import traceback
            traceback.print_exc()
            print(f'END {name} EXC {time.time() - t0:.1f}s', flush=True)
            msg = str(e).lower()
            if 'cuda' in msg or 'launch' in msg or 'illegal' in msg:
                sys.exit(3)   # context is probably poisoned
            continue
        print(f'END {name} {"OK" if ok else "FAIL"} {time.time() - t0:.1f}s', flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == '--worker':
        worker(sys.argv[2:])
        return
    names = [n for n in CASES if len(sys.argv) == 1 or any(n.startswith(a) for a in sys.argv[1:])]
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    log = open(os.path.join(ROOT, 'gpurun_out', 'debug_conv.log'), 'a')
    results = {}
    remaining = list(names)
    while remaining:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), '--worker'] + remaining,
                               capture_output=True, text=True, timeout=900)
            out = p.stdout + '\n--- stderr tail ---\n' + p.stderr[-4000:]
        except subprocess.TimeoutExpired as e:
            so = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or '')
            out = so + '\nTIMEOUT'
        print(out, flush=True)
        log.write(out + '\n')
        log.flush()
        began = None
        for line in out.splitlines():
            if line.startswith('BEGIN '):
                began = line.split()[1]
            elif line.startswith('END '):
                parts = line.split()
                results[parts[1]] = parts[2]
                began = None
        if began is not None:
            results[began] = 'CRASH'
        done = set(results)
        remaining = [n for n in remaining if n not in done]
        if began is None and remaining:
            # worker exited without starting the next case (poisoned context): just restart on the rest
            pass
    summ = 'SUMMARY ' + ' '.join(f'{n}:{results.get(n)}' for n in names)
    print(summ)
    log.write(summ + '\n')


if __name__ == '__main__':
    main()
