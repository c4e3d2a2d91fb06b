"""Frame path (one kernel per layer): pre-attention activation u in fp32 vs bf16 -- parity on a 270x480 input against the
CPU oracle and time per 1080p frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, recipe
from oracle import sr_torch_cpu
from rumpy_b200 import engine as E
from rumpy_b200.SISR.models.advanced.architectures import RCAN
dev = torch.device('cuda:0')
net = RCAN()
sd = {k: torch.from_numpy(v) for k, v in recipe.make_weights(recipe.rcan_spec(), seed=8).items()}
net.load_state_dict(sd); net = net.to(dev).eval()
x = torch.from_numpy(recipe.make_input((1, 3, 270, 480), seed=8))
with torch.no_grad():
    want = sr_torch_cpu.rcan_forward(sd, x, 10, 20, 4)
arch, kw = net._engine_kwargs()
xf = torch.rand((1, 3, 1080, 1920), device=dev)
for u_f32 in (True, False):
    kw2 = dict(kw); kw2['u_f32'] = u_f32
    eng = E.TrunkEngine(arch, list(net.parameters()), **kw2)
    with torch.no_grad():
        out = eng.forward(x.to(dev)).cpu()
        err = float((out - want).abs().max())
        eng.forward(xf); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): eng.forward(xf)
        e1.record(); e1.synchronize()
    print(f'u_f32={u_f32}: 270x480 max-abs vs CPU oracle {err:.5f}; 1080p frame {e0.elapsed_time(e1) / 3:.1f} ms', flush=True)
    del eng; torch.cuda.empty_cache()
