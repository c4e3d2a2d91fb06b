"""Bisects why bench.py's e2e loop is slower than tools/gpu_e2e_breakdown.py's stand-alone loop."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import recipe
import bench
from rumpy_b200.shared_framework.models import define_model

dev = torch.device('cuda:0')
h = define_model('rcan', device=0, model_save_dir=tempfile.mkdtemp(), eval_mode=True, scale=4)
h.net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.make_state_dict().items()}, strict=True)
h.net.eval()
eng = h.net.native_engine()
x_host = torch.from_numpy(recipe.make_input((16, 3, 48, 48), seed=8)).pin_memory()
x_dev = x_host.to(dev)


def e2e(tag, n=50):
    with torch.no_grad():
        for _ in range(3):
            h.run_eval(x_host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            h.run_eval(x_host)
        torch.cuda.synchronize()
        print('%-60s %.3f ms/call' % (tag, (time.perf_counter() - t0) / n * 1e3), flush=True)


e2e('fresh process, run_eval only')
with torch.no_grad():
    for _ in range(5):
        eng.forward_graphed(x_dev)
torch.cuda.synchronize()
e2e('after forward_graphed(x_dev) calls')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
e2e('after allocating the 256 MB flush buffer')
evs = []
with torch.no_grad():
    for _ in range(50):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.forward_graphed(x_dev); e1.record()
        evs.append((e0, e1))
torch.cuda.synchronize()
e2e('after the flushed device-timed loop')
s = bench.ClockSampler(0); s.start()
e2e('with the clock sampler thread')
s.result()
