"""Diagnostics behind tests/test_gpu_full_config.py: per-parameter gradient errors of the full RCAN train step and
the 1 000-step loss curves (b200 vs fp32 eager), dumped to gpurun_out/r02_full_parity.npz."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import recipe
from oracle import sr_torch_cpu
import test_gpu_full_config as T
from rumpy_b200 import train_native
from rumpy_b200.optim import FusedAdam
DEV = 'cuda:0'
what = sys.argv[1:] or ['grads', 'curve']
out = {}
if 'grads' in what:
    net, sd = T._net('rcan'); net.train()
    x = torch.from_numpy(recipe.make_input((16, 3, 64, 64), seed=80)); y = torch.from_numpy(recipe.make_input((16, 3, 256, 256), seed=180))
    eng = net.native_engine()
    o = eng.forward(x.to(DEV), training=True)
    loss, dy = train_native.l1_loss(o, y.to(DEV), want_grad=True)
    grads = eng.backward(x.to(DEV), dy)
    tr = sr_torch_cpu.Trainer(sd, 'rcan', lr=1e-4, n_resgroups=10, n_resblocks=20, scale=4)
    ref_loss, ref_out = tr.step(x, y)
    print('fwd max-abs', float((o.cpu() - ref_out).abs().max()), 'loss', loss.item(), ref_loss)
    ref = tr.grads()
    rows = []
    for (k, _), g in zip(net.named_parameters(), grads):
        r = ref[k].numpy().astype(np.float64); got = g.cpu().numpy().astype(np.float64)
        sc = max(np.abs(r).max(), 1e-30)
        cos = (got * r).sum() / (np.linalg.norm(got) * np.linalg.norm(r) + 1e-30)
        rows.append((k, np.abs(got - r).max() / sc, cos, sc, r.size))
    rows.sort(key=lambda t: -t[1])
    print('worst 25 by relative max error:')
    for k, e, c, sc, n in rows[:25]:
        print(f'  {k:40s} err {e:.4f} cos {c:.6f} |ref|max {sc:.3e} n {n}')
    rows.sort(key=lambda t: t[2])
    print('worst 15 by cosine:')
    for k, e, c, sc, n in rows[:15]:
        print(f'  {k:40s} err {e:.4f} cos {c:.6f} |ref|max {sc:.3e} n {n}')
    errs = np.array([r[1] for r in rows]); coss = np.array([r[2] for r in rows if r[4] >= 64])
    print('err percentiles 50/90/99/max', np.percentile(errs, [50, 90, 99, 100]), 'cos min/1%', coss.min(), np.percentile(coss, 1))
    kinds = {}
    for k, e, c, sc, n in rows:
        kind = 'bias' if k.endswith('bias') else ('ca' if 'conv_du' in k else 'w')
        kinds.setdefault(kind, []).append(e)
    print({k: (float(np.max(v)), float(np.median(v))) for k, v in kinds.items()})
    del net, eng, tr
    torch.cuda.empty_cache()
if 'curve' in what:
    lr = float(os.environ.get('LR', 1e-4))
    steps = int(os.environ.get('STEPS', 1000))
    net, sd = T._net('rcan'); net.train()
    opt = FusedAdam(list(net.parameters()), lr=lr)
    pairs = T._smooth_pairs(8, 4, 32, seed=8)
    kw = dict(n_resgroups=10, n_resblocks=20, scale=4)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    eager = sr_torch_cpu.Trainer({k: v.to(DEV) for k, v in sd.items()}, 'rcan', lr=lr, **kw)
    eager2 = sr_torch_cpu.Trainer({k: (v * (1 + 1e-6 * torch.randn_like(v))).to(DEV) for k, v in sd.items()}, 'rcan', lr=lr, **kw)
    ours, ref, ref2 = [], [], []
    for step in range(steps):
        x, y = pairs[step % len(pairs)]
        xd, yd = x.to(DEV), y.to(DEV)
        ours.append(train_native.train_step(net, opt, xd, yd)[0])
        ref.append(eager.step(xd, yd)[0])
        ref2.append(eager2.step(xd, yd)[0])      # the SAME fp32 oracle from weights perturbed by 1e-6: its own sensitivity
    ours = np.array([float(v) for v in ours]); ref = np.array(ref); ref2 = np.array(ref2)
    out.update(ours=ours, ref=ref, ref2=ref2)
    rel, rel2 = np.abs(ours - ref) / ref, np.abs(ref2 - ref) / ref
    for a in range(0, steps, 50):
        print(f'steps {a:4d}-{a+49:4d}: oracle {ref[a:a+50].mean():.4f} b200 {ours[a:a+50].mean():.4f}  max dev b200 {rel[a:a+50].max()*100:7.3f} %  '
              f'fp32-vs-fp32(1e-6 perturbed) {rel2[a:a+50].max()*100:7.3f} %')
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
np.savez(os.path.join(ROOT, 'gpurun_out', 'r02_full_parity.npz'), **out)
