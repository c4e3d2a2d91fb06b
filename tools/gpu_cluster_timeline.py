"""Per-layer timeline of the cluster kernel (clock64 stamps per CTA: 0 barrier A passed / layer start, 5 barrier B passed,
1 MMAs of the layer issued, 2 last accumulator complete, 4 pool exchanged, 3 last tile written)."""
import os, sys
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
eng.set_option('cluster_split', int(os.environ.get('SPLIT', 1)))
eng.set_option('cluster_dbg', int(os.environ.get('DBG', 0)))
with torch.no_grad():
    eng.forward(x)
    eng.set_timeline(dbg, LAYERS)
    for _ in range(2): eng.forward(x)
torch.cuda.synchronize()
print('mode', lib.rumpy_net_trunk_mode(eng.handle), 'split', eng.get_option('cluster_split'))
d = dbg.cpu()
for cta in (48,):
    t0 = d[cta, 0, 0].item()
    print(f'--- CTA {cta}')
    prev = t0
    for L in range(LAYERS):
        r = d[cta, L]
        rel = lambda k: (r[k].item() - r[0].item()) if r[k].item() else 0
        print(f'L{L:3d} start {r[0].item() - t0:8d} (+{r[0].item() - prev:6d})  B_ok +{rel(5):6d}  mma_issued +{rel(1):6d}  last_acc +{rel(2):6d}'
              f'  pool_pass_done +{rel(9):6d}  pool +{rel(4):6d}  y +{rel(6):6d}  apply0 +{rel(7):6d}  apply1 +{rel(8):6d}  last_tile_done +{rel(3):6d}')
        if r[4].item():
            d_ = lambda a, b: r[a].item() - r[b].item()
            print(f'       CA chain: acc->tmem_ld {d_(14, 2)}  transpose-sum {d_(15, 14)}  barrier+push+barrier {d_(9, 15)} | exchange wait {d_(4, 9)} | '
                  f'slot sums {d_(10, 4)}  barrier {d_(11, 10)}  FC+sigmoid {d_(12, 11)}  barrier {d_(13, 12)}  y regs {d_(6, 13)} | apply0 {d_(7, 6)} apply1 {d_(8, 7)}')
        prev = r[0].item()
eng.set_timeline(None)
