import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import torch, recipe
from rumpy_b200 import train_native
from rumpy_b200.optim import FusedAdam
from rumpy_b200.SISR.models.advanced.architectures import EDSR
dev = torch.device('cuda:0')
net = EDSR(net_features=256, num_blocks=32, res_scale=0.1).to(dev).train()
opt = FusedAdam(list(net.parameters()), lr=1e-4)
x = torch.rand((16, 3, 64, 64), device=dev); y = torch.rand((16, 3, 256, 256), device=dev)
for i in range(6):
    l, _ = train_native.train_step(net, opt, x, y)
    torch.cuda.synchronize()
    print('step', i, float(l), flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5): train_native.train_step(net, opt, x, y)
e1.record(); e1.synchronize()
print('edsr-full ms/step', e0.elapsed_time(e1) / 5)
