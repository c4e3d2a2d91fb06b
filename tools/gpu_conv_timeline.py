"""clock64 timeline of the per-layer conv kernel on a 1080p frame: the stamps of the LAST pooled conv (conv2 of the last RCAB;
conv_dbg bit 8 keeps the other convs from overwriting them).  Region 1: kernel-level stamps, region 2: the epilogue of
each CTA's third tile."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rumpy_b200 import engine as E
from rumpy_b200.SISR.models.advanced.architectures import RCAN
dev = torch.device('cuda:0')
net = RCAN(n_resgroups=1, n_resblocks=2).to(dev).eval()
arch, kw = net._engine_kwargs()
xf = torch.rand((1, 3, 1080, 1920), device=dev)
eng = E.TrunkEngine(arch, list(net.parameters()), **kw)
eng.set_option('conv_dbg', 8 | int(os.environ.get('MODE', 0)))
GRID = 148
dbg = torch.zeros((2 * GRID, 16), dtype=torch.int64, device=dev)
eng.set_timeline(dbg, 1)
with torch.no_grad():
    for _ in range(2): eng.forward(xf)
torch.cuda.synchronize()
d = dbg.cpu()
tiles = (1080 // 8) * (1920 // 16)
per_cta = tiles / GRID
for cta in (0, 1, 73, 147):
    r, e = d[cta], d[GRID + cta]
    rel = lambda k: int(r[k] - r[0])
    print(f'CTA {cta}: prologue {rel(1)}  first full->MMA {rel(4)}  tile0 MMAs issued {rel(5)}  tile0 acc ready {rel(6)} staged {rel(8)}'
          f' | last tile acc ready {rel(9)} staged {rel(10)} | loop end {rel(11)} stores done {rel(12)} exit {rel(13)}'
          f' | cycles/tile {rel(11) / per_cta:.0f}  ns total {int(r[15] - r[14])}')
    b = int(e[0])
    print('    tile 40 epilogue: wait acc %d | tmem_ld %d | math+sts %d | fence+storewait %d | barrier %d | tma store %d | pool %d'
          % tuple(int(e[k + 1] - e[k]) for k in range(7)))
    print('    waits over the whole kernel: MMA warp on accumulator %d, on A stages %d; producer on free stages %d (cycles, of %d)'
          % (int(e[12]), int(e[13]), int(e[14]), rel(13)))
    print('    tile 40, relative to the MMA warp being ready for it: accumulator free +%d, MMAs issued +%d, epilogue sees the accumulator +%d, epilogue was ready for it at +%d, epilogue done +%d'
          % (int(e[11] - e[10]), int(e[15] - e[10]), int(e[1] - e[10]), int(e[0] - e[10]), int(e[7] - e[10])))
