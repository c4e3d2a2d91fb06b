"""One eager (non-graph) RCAN cfg#2 forward between cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import recipe
from rumpy_b200.SISR.models.advanced.architectures import RCAN

dev = torch.device('cuda:0')
net = RCAN()
net.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(recipe.rcan_spec(), seed=8).items()})
net = net.to(dev).eval()
x = torch.from_numpy(recipe.make_input((16, 3, 48, 48), seed=8)).to(dev)
eng = net.native_engine()
with torch.no_grad():
    eng.forward(x); eng.forward(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eng.forward(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print('profiled one forward,', eng.lib.rumpy_net_num_launches(eng.handle), 'launches')
