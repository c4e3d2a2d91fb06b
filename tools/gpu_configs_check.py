"""BASELINE.json configs[3] (EDSR-full training) and configs[4] (RCAN 1080p frame inference): do they run, are they
right (vs CPU oracle on a crop-sized problem), how fast."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import recipe
from oracle import sr_torch_cpu
from rumpy_b200.SISR.models.advanced.architectures import RCAN, EDSR
from rumpy_b200 import train_native
from rumpy_b200.optim import FusedAdam
dev = torch.device('cuda:0')

def ev(): return torch.cuda.Event(enable_timing=True)

which = sys.argv[1:] or ['edsrfull_train', 'frame']
if 'edsrfull_train' in which:
    spec = recipe.edsr_spec(32, 256, 4)
    sd = recipe.make_weights(spec, seed=8)
    net = EDSR(net_features=256, num_blocks=32, res_scale=0.1)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); net = net.to(dev).train()
    opt = FusedAdam(list(net.parameters()), lr=1e-4)
    # parity on a small batch vs CPU oracle (3 steps)
    x = recipe.make_input((2, 3, 24, 24), 1); y = recipe.make_input((2, 3, 96, 96), 2)
    tr = sr_torch_cpu.Trainer({k: torch.from_numpy(v) for k, v in sd.items()}, 'edsr', lr=1e-4, num_blocks=32, res_scale=0.1, scale=4)
    for i in range(3):
        lg = train_native.train_step(net, opt, torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev))[0].item()
        lc, _ = tr.step(torch.from_numpy(x), torch.from_numpy(y))
        print(f'[EDSR-full train] step {i}: loss gpu {lg:.6f} cpu {lc:.6f} rel {abs(lg-lc)/lc:.2e}', flush=True)
    # throughput at 16 x 64 x 64
    x = torch.from_numpy(recipe.make_input((16, 3, 64, 64), 3)).to(dev); y = torch.from_numpy(recipe.make_input((16, 3, 256, 256), 4)).to(dev)
    for _ in range(3): train_native.train_step(net, opt, x, y)
    torch.cuda.synchronize(); e0, e1 = ev(), ev(); e0.record()
    for _ in range(5): train_native.train_step(net, opt, x, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 3 * 100505088 * 16 * 64 * 64
    print(f'[EDSR-full train 16x64x64] {ms:.2f} ms/step, {16/ms*1e3:.1f} patches/s, {fl/ms*1e-9:.1f} TFLOP/s, '
          f'mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB', flush=True)
    del net, opt; torch.cuda.empty_cache()
if 'frame' in which:
    sd = recipe.make_weights(recipe.rcan_spec(), seed=8)
    net = RCAN(); net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); net = net.to(dev).eval()
    eng = net.native_engine()
    with torch.no_grad():
        # parity on a quarter-resolution frame (CPU oracle on 1080p takes minutes)
        x = recipe.make_input((1, 3, 135, 240), 5)
        out = eng.forward(torch.from_numpy(x).to(dev)).cpu().numpy()
        torch.set_num_threads(os.cpu_count())
        ref = sr_torch_cpu.rcan_forward({k: torch.from_numpy(v) for k, v in sd.items()}, torch.from_numpy(x)).numpy()
        print(f'[RCAN frame 135x240] max-abs err {np.abs(out-ref).max():.5f}', flush=True)
        x = torch.from_numpy(recipe.make_input((1, 3, 1080, 1920), 6)).to(dev)
        for _ in range(2): eng.forward(x)
        torch.cuda.synchronize(); e0, e1 = ev(), ev(); e0.record()
        for _ in range(3): o = eng.forward(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f'[RCAN frame 1080x1920 -> 4320x7680] {ms:.1f} ms/frame, {4320*7680/ms*1e-3:.1f} Mpix/s, '
              f'{66.01e12/ms*1e-9:.0f} TFLOP/s, finite={bool(torch.isfinite(o).all())}, mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB', flush=True)
