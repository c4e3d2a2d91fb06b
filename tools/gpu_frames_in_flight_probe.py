"""1080p frames: one at a time vs two in flight (parallel.FramesInFlight), ms per frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, recipe
from rumpy_b200 import parallel
from rumpy_b200.SISR.models.advanced.architectures import RCAN
dev = torch.device('cuda:0')
net = RCAN()
net.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(recipe.rcan_spec(), seed=8).items()})
net = net.to(dev).eval()
H, W = int(os.environ.get('H', 1080)), int(os.environ.get('W', 1920))
frames = [torch.rand((1, 3, H, W), device=dev) for _ in range(4)]
for depth in (1, 2, 1, 2):
    pipe = parallel.FramesInFlight(net, depth=depth)
    pipe.run(frames[:2], consume=lambda i, o: None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pipe.run(frames, consume=lambda i, o: None); e1.record(); e1.synchronize()
    print(f'depth {depth}: {e0.elapsed_time(e1) / 4:.1f} ms per frame', flush=True)
    del pipe; torch.cuda.empty_cache()
