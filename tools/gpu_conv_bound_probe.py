"""What bounds the per-layer conv kernel at 1080p: frame time with parts of the kernel switched off (ConvArgs::dbg_mode:
1 = A boxes fetched only for the first fills, 4 = no MMAs, 16 = epilogue stages nothing, 32 = no pool sums).  Results are garbage in modes != 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rumpy_b200 import engine as E
from rumpy_b200.SISR.models.advanced.architectures import RCAN
dev = torch.device('cuda:0')
net = RCAN().to(dev).eval()
arch, kw = net._engine_kwargs()
xf = torch.rand((1, 3, 1080, 1920), device=dev)
for mode in [int(v) for v in os.environ.get('MODES', '0,1,2,4,3,5,6,7').split(',')]:
    eng = E.TrunkEngine(arch, list(net.parameters()), **kw)
    if mode:
        eng.set_option('conv_dbg', mode)
    with torch.no_grad():
        eng.forward(xf); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): eng.forward(xf)
        e1.record(); e1.synchronize()
    print(f'conv_dbg={mode} (no-A-fetch {mode & 1}, no-MMA {(mode >> 2) & 1}, no-staging {(mode >> 4) & 1}, no-pool {(mode >> 5) & 1}, no-stores {(mode >> 6) & 1}): 1080p frame {e0.elapsed_time(e1) / 3:.1f} ms', flush=True)
    del eng; torch.cuda.empty_cache()
