"""Backward / train-step probe on the GPU box: per-parameter gradient error vs the reference golden vectors,
3 Adam steps vs golden losses."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import recipe
from rumpy_b200.SISR.models.advanced.architectures import RCAN, EDSR
from rumpy_b200 import train_native
from rumpy_b200.optim import FusedAdam

dev = torch.device('cuda:0')

def build(arch, kw, sd):
    if arch == 'rcan':
        net = RCAN(n_resblocks=kw['n_resblocks'], n_resgroups=kw['n_resgroups'], n_feats=kw['n_feats'], scale=kw['scale'])
    else:
        net = EDSR(net_features=kw['n_feats'], num_blocks=kw['num_blocks'], scale=kw['scale'], res_scale=kw['res_scale'])
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return net.to(dev).train()

for name in (sys.argv[1:] or list(recipe.CASES)):
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
    arch, kw, sd, x, y = recipe.case_tensors(name)
    net = build(arch, kw, sd)
    eng = net.native_engine()
    xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    out = eng.forward(xt, training=True)
    loss, dy = train_native.l1_loss(out, yt, want_grad=True)
    grads = eng.backward(xt, dy)
    torch.cuda.synchronize()
    print(f'== {name}: loss {loss.item():.6f} (ref {float(gold["loss"]):.6f}); fwd max-abs {np.abs(out.cpu().numpy()-gold["out"]).max():.5f}')
    worst = []
    for (k, _), g in zip(net.named_parameters(), grads):
        ref = gold['gradsub::' + k]
        got = recipe.subsample(g.cpu().numpy())
        scale = max(np.abs(ref).max(), 1e-12)
        rel = np.abs(got - ref).max() / scale
        cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
        worst.append((rel, cos, k, scale))
    worst.sort(reverse=True)
    for rel, cos, k, scale in worst[:12]:
        print(f'   rel_err {rel:8.4f}  cos {cos:7.4f}  ref_absmax {scale:.3e}  {k}')
    print(f'   median rel_err {np.median([w[0] for w in worst]):.4f}, min cos {min(w[1] for w in worst):.4f}')
    # 3 optimiser steps
    net = build(arch, kw, sd)
    opt = FusedAdam(list(net.parameters()), lr=1e-4)
    losses = []
    for _ in range(3):
        l, _o = train_native.train_step(net, opt, xt, yt)
        losses.append(l.item())
    print('   train losses', ['%.6f' % v for v in losses], 'ref', ['%.6f' % v for v in gold['train_losses']])
