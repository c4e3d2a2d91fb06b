"""One eager pass of every widening path between cudaProfilerStart/Stop, for
ncu --profile-from-start off:  Q-RCAN forward (16x48x48) and train step (16x64x64), HAN forward (16x48x48),
eval glue on a x4 1080p frame, one device patch batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np
import torch
import recipe
from rumpy_b200 import train_native
from rumpy_b200.optim import FusedAdam
from rumpy_b200.SISR.models.advanced.architectures import HAN
from rumpy_b200.SISR.models.attention_manipulators.architectures import QRCAN
from rumpy_b200.shared_framework.data import DevicePairSet, psnr_y_device, quantize_u8_device

dev = torch.device('cuda:0')


def load(net, seed=8, fix=None):
    spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    sd = recipe.make_weights(spec, seed=seed)
    sd.update(fix or {})
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return net.to(dev)


qnet = load(QRCAN(style='standard', num_metadata=10, include_q_layer=True)).eval()
hnet = load(HAN(), fix={'la.gamma': np.array([0.3], np.float32), 'csa.gamma': np.array([0.5], np.float32)}).eval()
tnet = load(QRCAN(style='standard', num_metadata=10, include_q_layer=True)).train()
opt = FusedAdam(list(tnet.parameters()), lr=1e-4)
x = torch.rand((16, 3, 48, 48), device=dev)
meta = torch.rand((16, 10, 1, 1), device=dev)
xt, yt = torch.rand((16, 3, 64, 64), device=dev), torch.rand((16, 3, 256, 256), device=dev)
fsr, fhr = torch.rand((1, 3, 4320, 7680), device=dev), torch.rand((1, 3, 4320, 7680), device=dev)
ds = DevicePairSet({'synthetic': 16, 'crop': 64, 'random_augment': True}, 4, seed=8, device=0)


def one_pass():
    with torch.no_grad():
        qnet.native_engine().set_metadata(meta, 16)
        qnet.native_engine().forward(x)
        hnet.native_engine().forward(x)
    train_native.train_step(tnet, opt, xt, yt, metadata=meta)
    psnr_y_device(fsr, fhr)
    quantize_u8_device(fsr)
    next(iter(ds.batches(16)))


for _ in range(2):
    one_pass()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
one_pass()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('profiled one pass of the widening paths')
