"""Train-step time of the full RCAN (16x64x64) for different split-K granularities of the batched wgrad kernel."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import recipe
from rumpy_b200 import _lib, train_native
from rumpy_b200.optim import FusedAdam
from rumpy_b200.SISR.models.advanced.architectures import RCAN

lib = _lib.load()
lib.rumpy_debug_set_wgrad_split.argtypes = [ctypes.c_int]
dev = torch.device('cuda:0')
x, y = torch.rand((16, 3, 64, 64), device=dev), torch.rand((16, 3, 256, 256), device=dev)
for tiles in (32, 64, 128, 256, 32):
    lib.rumpy_debug_set_wgrad_split(tiles)
    net = RCAN()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(recipe.rcan_spec(), seed=8).items()})
    net = net.to(dev).train()
    opt = FusedAdam(list(net.parameters()), lr=1e-4)
    for _ in range(5):
        train_native.train_step(net, opt, x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        loss, _ = train_native.train_step(net, opt, x, y)
    e1.record(); e1.synchronize()
    print(f'tiles_per_split {tiles}: {e0.elapsed_time(e1) / 20:.3f} ms/step, loss {loss.item():.5f}', flush=True)
    del net, opt
    torch.cuda.empty_cache()
