"""Depth sweep of the role-swapped band kernel against the per-layer path (GPU box)."""
import ctypes, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden'))
import numpy as np, torch
import recipe
from rumpy_b200 import _lib
from rumpy_b200.SISR.models.advanced.architectures import RCAN

lib = _lib.load()
dev = torch.device('cuda:0')
shape = tuple(int(v) for v in os.environ.get('SHAPE', '16,3,48,48').split(','))
for g, b in [tuple(int(v) for v in t.split("x")) for t in os.environ.get("NETS", "1x3,1x20,2x20,10x20").split(",")]:
    net = RCAN(n_resgroups=g, n_resblocks=b)
    spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    sd = recipe.make_weights(spec, seed=8)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    net = net.to(dev).eval()
    x = torch.from_numpy(recipe.make_input(shape, seed=8)).to(dev)
    outs = {}
    for name, (trunk, band) in {'per-layer': (0, 0), 'band': (1, 1), 'band2': (1, 1), 'cluster': (1, 0)}.items():
        eng = net.native_engine(); eng.set_option('trunk', trunk); eng.set_option('band', band)
        with torch.no_grad():
            outs[name] = eng.forward(x).clone()
        torch.cuda.synchronize()
        mode = lib.rumpy_net_trunk_mode(eng.handle)
        outs[name + '_mode'] = mode
    base = outs['per-layer']
    def stat(o):
        d = (o - base).abs()
        return f'max {float(d.max()):.4g} nan {int(torch.isnan(o).sum())} imgs_bad {[int(i) for i in torch.nonzero(torch.isnan(o).flatten(1).any(1)).flatten()][:16]}'
    print(f'{g}x{b}: band(mode {outs["band_mode"]}) {stat(outs["band"])} | rerun equal {bool((outs["band"] == outs["band2"]).all())} | cluster(mode {outs["cluster_mode"]}) {stat(outs["cluster"])}', flush=True)
