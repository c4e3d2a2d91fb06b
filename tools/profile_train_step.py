"""One eager RCAN cfg#3 train step (16x64x64) between cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import recipe
from rumpy_b200.SISR.models.advanced.architectures import RCAN
from rumpy_b200 import train_native
from rumpy_b200.optim import FusedAdam

dev = torch.device('cuda:0')
net = RCAN()
net.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(recipe.rcan_spec(), seed=8).items()})
net = net.to(dev).train()
opt = FusedAdam(list(net.parameters()), lr=1e-4)
x = torch.from_numpy(recipe.make_input((16, 3, 64, 64), seed=8)).to(dev)
y = torch.from_numpy(recipe.make_input((16, 3, 256, 256), seed=9)).to(dev)
for _ in range(2):
    train_native.train_step(net, opt, x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
train_native.train_step(net, opt, x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('profiled one train step')
