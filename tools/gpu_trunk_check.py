"""Persistent trunk kernel (trunk_pipe.cuh) vs the per-layer path: agreement, determinism and timing (GPU box).
python tools/gpu_trunk_check.py [small] [rcan2] [rcan3] [edsr] [timeline]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import recipe
from rumpy_b200 import _lib
from rumpy_b200.SISR.models.advanced.architectures import RCAN, EDSR

dev = torch.device('cuda:0')
lib = _lib.load()


def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def build(kind, **kw):
    net = RCAN(**kw) if kind == 'rcan' else EDSR(**kw)
    spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    sd = recipe.make_weights(spec, seed=8)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return net.to(dev).eval()


def compare(name, net, xshape, time_it=False, flop_per_px=31835520):
    x = torch.from_numpy(recipe.make_input(xshape, seed=8)).to(dev)
    outs = {}
    for mode, (trunk, cluster) in (('per-layer', (0, 0)), ('trunk-flags', (1, 0)), ('trunk-cluster', (1, 1))):
        eng = net.native_engine()
        eng.set_option('trunk', trunk); eng.set_option('cluster', cluster)
        with torch.no_grad():
            o1 = eng.forward(x).clone()
            o2 = eng.forward(x).clone()
            torch.cuda.synchronize()
            outs[mode] = o1
            det = bool((o1 == o2).all())
            launches = lib.rumpy_net_num_launches(eng.handle)
            msg = f'[{name}] {mode}: launches/forward {launches}, deterministic {det}'
            if time_it:
                ms = timeit(lambda: eng.forward_graphed(x))
                px = xshape[0] * xshape[2] * xshape[3]
                msg += f', graph {ms:.3f} ms -> {flop_per_px * px / ms * 1e-9:.1f} TFLOP/s, {px * 16 / ms * 1e-3:.1f} Mpix/s out'
            print(msg, flush=True)
    for mode in ('trunk-flags', 'trunk-cluster'):
        d = (outs['per-layer'] - outs[mode]).abs().max().item()
        print(f'[{name}] max |{mode} - per-layer| = {d:.3e} (out absmax {outs["per-layer"].abs().max().item():.3f})', flush=True)
    net.native_engine().set_option('trunk', 1); net.native_engine().set_option('cluster', 1)


def timeline(net, xshape, layers=12):
    """Per-CTA clock64 stamps of the first `layers` trunk layers (slot meaning: trunk_pipe.cuh TR_STAMP)."""
    x = torch.from_numpy(recipe.make_input(xshape, seed=8)).to(dev)
    eng = net.native_engine()
    with torch.no_grad():
        eng.forward(x)
    buf = torch.zeros(148 * layers * 2 * 16, dtype=torch.int64, device=dev)
    eng.set_timeline(buf, layers)
    with torch.no_grad():
        eng.forward(x)
    torch.cuda.synchronize()
    eng.set_timeline(None)
    t = buf.cpu().numpy().reshape(148, layers, 2, 16)
    names = ['deps_ok', 'mma_start', 'mma_commit', 'epi_start', 'pool_done', 'y_ready', 'published', '-', 'bias_bar', 'staged',
             'store_go', 'store_done', 'pool_red', 'cnt_seen', 'y_seen', 'mean_ok']
    for cta in (0,):
        base = t[cta, 0, 0, 0]
        print(f'--- CTA {cta}: cycles since layer-0 tile-0 deps_ok')
        for L in range(layers):
            for j in range(2):
                row = ' '.join(f'{names[s]}={int(t[cta, L, j, s] - base) if t[cta, L, j, s] else -1:>7}' for s in (0, 1, 2, 3, 8, 12, 4, 13, 15, 5, 14, 9, 10, 11, 6)) + f' passes={int(t[cta, L, j, 7])}'
                print(f'  L{L:02d} j{j}: {row}')
    per_layer = (t[:, layers - 1, 0, 1] - t[:, 1, 0, 1]) / float(layers - 2)
    print(f'median cycles per layer (mma_start to mma_start, tile 0): {np.median(per_layer[per_layer > 0]):.0f}')


if __name__ == '__main__':
    which = sys.argv[1:] or ['small', 'rcan2']
    sync_mode, no_cluster = None, False
    if 'small' in which:
        compare('RCAN 1g2b 2x16x16', build('rcan', n_resgroups=1, n_resblocks=2), (2, 3, 16, 16))
        compare('RCAN 2g3b 3x20x37 ragged', build('rcan', n_resgroups=2, n_resblocks=3), (3, 3, 20, 37))
        compare('RCAN 2g2b 1x64x96 one image, 4 slots', build('rcan', n_resgroups=2, n_resblocks=2), (5, 3, 64, 96))
        compare('RCAN 1g2b 1x100x200 one image over 2 slots (pool-all-then-apply order)',
                build('rcan', n_resgroups=1, n_resblocks=2), (1, 3, 100, 200))
        compare('EDSR 4 blocks 2x24x24', build('edsr', num_blocks=4), (2, 3, 24, 24))
    if 'rcan2' in which:
        compare('RCAN cfg2 16x48x48', build('rcan'), (16, 3, 48, 48), time_it=True)
    if 'rcan3' in which:
        compare('RCAN 16x64x64', build('rcan'), (16, 3, 64, 64), time_it=True)
    if 'edsr' in which:
        compare('EDSR-baseline 16x48x48', build('edsr'), (16, 3, 48, 48), time_it=True, flop_per_px=3966336)
    if 'syncmodes' in which:
        net = build('rcan'); x = torch.from_numpy(recipe.make_input((16, 3, 48, 48), seed=8)).to(dev)
        eng = net.native_engine()
        eng.set_option('trunk', 0)
        with torch.no_grad():
            ref = eng.forward(x).clone()
        eng.set_option('trunk', 1)
        for mode in (15, 0, 8, 4, 12, 1, 2, 3, 9):
            eng.set_option('trunk_sync_mode', mode)
            with torch.no_grad():
                outs = [eng.forward(x).clone() for _ in range(6)]
                torch.cuda.synchronize()
                ms = timeit(lambda: eng.forward(x), iters=10, warm=2)
            det = all(bool((o == outs[0]).all()) for o in outs)
            print(f'sync_mode {mode:2d}: deterministic {det}, max diff vs per-layer '
                  f'{max((o - ref).abs().max().item() for o in outs):.3e}, eager {ms:.3f} ms', flush=True)
        eng.set_option('trunk_sync_mode', 8)
    for w in which:
        if w.startswith('sync='):
            sync_mode = int(w[5:]); print('sync_mode', w[5:])
    if 'ctimeline' in which:
        net = build('rcan'); x = torch.from_numpy(recipe.make_input((16, 3, 48, 48), seed=8)).to(dev)
        eng = net.native_engine()
        with torch.no_grad():
            eng.forward(x)
        layers = 10
        buf = torch.zeros(148 * layers * 16, dtype=torch.int64, device=dev)
        eng.set_timeline(buf, layers)
        with torch.no_grad():
            eng.forward(x)
        torch.cuda.synchronize()
        eng.set_timeline(None)
        t = buf.cpu().numpy().reshape(148, layers, 16)
        names = ['in_full', 'mma_issued', 'last_acc', 'last_tile_done', 'pool_full', 'in_full_b']
        for cta in (0, 1, 5):
            base = t[cta, 0, 0]
            print(f'--- cluster CTA {cta}: cycles since layer-0 in_full')
            for L in range(layers):
                print(f'  L{L:02d}: ' + ' '.join(f'{names[s]}={int(t[cta, L, s] - base) if t[cta, L, s] else -1:>7}' for s in range(6)))
    if 'timeline' in which:
        timeline(build('rcan'), (16, 3, 48, 48), layers=6)
    if 'timeline3' in which:
        net3 = build('rcan'); net3.native_engine().set_option('cluster', 0)
        timeline(net3, (16, 3, 64, 64), layers=8)
    if 'time2' in which:
        net = build('rcan'); x = torch.from_numpy(recipe.make_input((16, 3, 48, 48), seed=8)).to(dev)
        with torch.no_grad():
            print('cfg2 graph ms', timeit(lambda: net.native_engine().forward_graphed(x)))
