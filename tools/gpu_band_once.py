"""One forward of a G x B RCAN at 16x48x48 through the default plan (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rumpy_b200.SISR.models.advanced.architectures import RCAN
dev = torch.device('cuda:0')
G, B = int(os.environ.get('G', 2)), int(os.environ.get('B', 20))
net = RCAN(n_resgroups=G, n_resblocks=B).to(dev).eval()
x = torch.rand((16, 3, 48, 48), device=dev)
eng = net.native_engine()
eng.set_option('band', int(os.environ.get('BAND', 1)))
with torch.no_grad():
    for _ in range(int(os.environ.get('REPS', 3))):
        eng.forward(x)
torch.cuda.synchronize()
