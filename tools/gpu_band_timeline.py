"""Per-layer timeline of the band kernel (clock64 stamps: 0 MMA start, 1 MMA issue done, 2 last chunk's accumulator read /
CA pass 1 done, 4 pool ready, 3 epilogue done), first CTA of a few clusters."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, recipe
from rumpy_b200 import _lib
from rumpy_b200.SISR.models.advanced.architectures import RCAN
lib = _lib.load()
dev = torch.device('cuda:0')
G, B = int(os.environ.get('G', 2)), int(os.environ.get('B', 4))
net = RCAN(n_resgroups=G, n_resblocks=B).to(dev).eval()
x = torch.rand((16, 3, 48, 48), device=dev)
LAYERS = G * (2 * B + 1) + 1
dbg = torch.zeros((96, LAYERS, 16), dtype=torch.int64, device=dev)
eng = net.native_engine()
eng.set_option('band', int(os.environ.get('BAND', 1)))
eng.set_timeline(dbg, LAYERS)
with torch.no_grad():
    for _ in range(3): eng.forward(x)
torch.cuda.synchronize()
print('mode', lib.rumpy_net_trunk_mode(eng.handle))
d = dbg.cpu()
for cta in (48, 53):
    t0 = d[cta, 0, 0].item()
    print(f'--- CTA {cta}')
    prev = t0
    for L in range(LAYERS):
        r = d[cta, L]
        rel = lambda k: (r[k].item() - r[0].item()) if r[k].item() else 0
        print(f'L{L:3d} start {r[0].item() - t0:8d} (+{r[0].item() - prev:6d})  mma_issued +{rel(1):6d}  acc_read +{rel(2):6d}'
              f'  pool +{rel(4):6d}  epi_done +{rel(3):6d} | acc_empty ok {rel(8):5d} {rel(9):5d} {rel(10):5d} | chunk issued {rel(5):5d} {rel(6):5d} {rel(7):5d}'
              f' | epi loaded {rel(11):5d} {rel(12):5d} {rel(13):5d} | cp0 {rel(14):5d} {rel(15):5d}')
        prev = r[0].item()
eng.set_timeline(None)
