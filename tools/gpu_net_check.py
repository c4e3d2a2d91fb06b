"""Quick whole-network accuracy + timing probe (GPU box).  python tools/gpu_net_check.py [u_bf16]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import recipe
from oracle import sr_torch_cpu
from rumpy_b200.SISR.models.advanced.architectures import RCAN, EDSR

dev = torch.device('cuda:0')

def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def run(name, net, spec, xshape, flop_per_px, check=True):
    sd = recipe.make_weights(spec, seed=8)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    net = net.to(dev).eval()
    x = recipe.make_input(xshape, seed=8)
    xt = torch.from_numpy(x).to(dev)
    with torch.no_grad():
        out = net(xt)
        if check:
            torch.set_num_threads(os.cpu_count())
            t0 = time.time()
            arch, kw = sr_torch_cpu.infer_arch({k: torch.from_numpy(v) for k, v in sd.items()})
            ref = sr_torch_cpu.forward({k: torch.from_numpy(v) for k, v in sd.items()}, torch.from_numpy(x), arch,
                                       res_scale=0.1, **kw).numpy()
            cpu_s = time.time() - t0
            err = np.abs(out.cpu().numpy() - ref).max()
            print(f'[{name}] max-abs err vs CPU oracle = {err:.5f} (ref absmax {np.abs(ref).max():.3f}), cpu {cpu_s:.2f}s', flush=True)
        eng = net.native_engine()
        ms = timeit(lambda: eng.forward(xt))
        msg = timeit(lambda: eng.forward_graphed(xt))
    px = xshape[0] * xshape[2] * xshape[3]
    fl = flop_per_px * px
    s = 4
    print(f'[{name}] eager-launch {ms:.3f} ms, graph {msg:.3f} ms -> {fl/msg*1e-9:.1f} TFLOP/s, '
          f'{px*s*s/msg*1e-3:.1f} Mpix/s out', flush=True)

if __name__ == '__main__':
    which = sys.argv[1:] or ['rcan2', 'edsr', 'rcan3']
    if 'rcan2' in which:
        run('RCAN cfg2 16x48x48', RCAN(), recipe.rcan_spec(), (16, 3, 48, 48), 31835520)
    if 'rcan3' in which:
        run('RCAN 16x64x64', RCAN(), recipe.rcan_spec(), (16, 3, 64, 64), 31835520, check=False)
    if 'edsr' in which:
        run('EDSR-baseline 16x48x48', EDSR(), recipe.edsr_spec(), (16, 3, 48, 48), 3966336)
    if 'edsrfull' in which:
        run('EDSR-full 4x48x48', EDSR(net_features=256, num_blocks=32, res_scale=0.1), recipe.edsr_spec(32, 256), (4, 3, 48, 48), 100505088)
    if 'frame' in which:
        run('RCAN 1x270x480', RCAN(), recipe.rcan_spec(), (1, 3, 270, 480), 31835520, check=False)
