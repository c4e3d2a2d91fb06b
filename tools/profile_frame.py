"""One eager RCAN x4 forward of a whole 1920x1080 frame (BASELINE configs[4]) between cudaProfilerStart/Stop, for
ncu --profile-from-start off.  H, W from the environment (defaults 1080, 1920)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import recipe
from rumpy_b200.SISR.models.advanced.architectures import RCAN

dev = torch.device('cuda:0')
H, W = int(os.environ.get('H', 1080)), int(os.environ.get('W', 1920))
net = RCAN()
net.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(recipe.rcan_spec(), seed=8).items()})
net = net.to(dev).eval()
x = torch.rand((1, 3, H, W), device=dev)
eng = net.native_engine()
with torch.no_grad():
    eng.forward(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eng.forward(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print('profiled one frame forward,', eng.lib.rumpy_net_num_launches(eng.handle), 'launches, trunk mode',
      eng.lib.rumpy_net_trunk_mode(eng.handle))
