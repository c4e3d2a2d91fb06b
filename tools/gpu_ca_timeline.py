import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, recipe
from rumpy_b200 import _lib
from rumpy_b200.SISR.models.advanced.architectures import RCAN
lib = _lib.load()
lib.rumpy_debug_set_timeline.argtypes = [ctypes.c_void_p]
dev = torch.device('cuda:0')
net = RCAN(n_resgroups=1, n_resblocks=1).to(dev).eval()
x = torch.rand((16, 3, 48, 48), device=dev)
dbg = torch.zeros((2 * 148, 16), dtype=torch.int64, device=dev)
lib.rumpy_debug_set_timeline(dbg.data_ptr())     # baked into the plan at build time
eng = net.native_engine()
with torch.no_grad():
    for _ in range(3): eng.forward(x)
torch.cuda.synchronize()
d = dbg.cpu()[148:]
names = {0: 'entry', 1: 'phase1 done', 2: 'barrier passed', 3: 'y ready', 4: 'tile0 staged(p2)', 5: 'tile1 staged(p2)', 12: 'done'}
rel = d - d[:, :1]
for s in sorted(names):
    col = rel[:, s]; col = col[col > 0]
    print(f'{names[s]:20s} cta0 {rel[0, s].item():8d}  cta73 {rel[73, s].item():8d}  median {col.median().item() if len(col) else 0:8.0f}  max {col.max().item() if len(col) else 0:8.0f}')
