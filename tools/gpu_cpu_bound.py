"""Is the train step CPU-launch-bound?  Host time to ENQUEUE a step vs device time per step."""
import os, sys, time
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
eng = net.native_engine()
for _ in range(3):
    train_native.train_step(net, opt, x, y)
torch.cuda.synchronize()
# host enqueue time, device kept far behind by queueing 5 steps
t0 = time.perf_counter()
for _ in range(5):
    train_native.train_step(net, opt, x, y)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host enqueue per step {(t1-t0)/5*1e3:.2f} ms; total per step incl. drain {(t2-t0)/5*1e3:.2f} ms')
# phase split with events
def ev(): return torch.cuda.Event(enable_timing=True)
es = [ev() for _ in range(6)]
es[0].record()
out = eng.forward(x, training=True); es[1].record()
loss, dy = train_native.l1_loss(out, y, want_grad=True); es[2].record()
eng.backward(x, dy); es[3].record()
opt.step(); es[4].record()
eng.refresh_weights(training=True); es[5].record()
torch.cuda.synchronize()
names = ['forward', 'l1', 'backward', 'adam', 'repack']
print(' | '.join(f'{n} {es[i].elapsed_time(es[i+1]):.2f} ms' for i, n in enumerate(names)))

# whole train step in a CUDA graph (lr / Adam step count baked in: timing experiment only)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    train_native.train_step(net, opt, x, y)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    train_native.train_step(net, opt, x, y)
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = ev(), ev()
e0.record()
for _ in range(10): g.replay()
e1.record(); torch.cuda.synchronize()
print(f'graph replay of the full train step: {e0.elapsed_time(e1)/10:.2f} ms per step')
