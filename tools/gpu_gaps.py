"""Back-to-back conv launches inside one CUDA graph: per-kernel start/end in globaltimer ns -> launch gaps."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rumpy_b200 import ops, _lib
lib = _lib.load()
lib.rumpy_debug_set_timeline.argtypes = [ctypes.c_void_p]
dev = torch.device('cuda:0')
N, H, W, C = 16, 48, 48, 64
x = torch.rand((N, H, W, C), device=dev).to(torch.bfloat16)
w = (torch.rand((C, C, 3, 3), device=dev) - 0.5) / 24
b = torch.rand((C,), device=dev)
wp = ops.pack_conv3x3(w)
ys = [torch.empty_like(x) for _ in range(2)]
K = 8
dbg = torch.zeros((K, 148, 16), dtype=torch.int64, device=dev)
def seq(with_dbg):
    src = x
    for i in range(K):
        lib.rumpy_debug_set_timeline(dbg[i].data_ptr() if with_dbg else None)
        ops.conv3x3(src, wp, b, out_bf16=ys[i & 1], N=N, H=H, W=W, Cin=C, Cout=C, relu=True)
        src = ys[i & 1]
    lib.rumpy_debug_set_timeline(None)
for _ in range(3): seq(False)
torch.cuda.synchronize()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): seq(True)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): seq(True)
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f'graph of {K} convs: {e0.elapsed_time(e1)*1e3/K:.2f} us per conv')
d = dbg.cpu()
start = d[:, :, 14]; end = d[:, :, 15]
t0 = start[0].min().item()
for i in range(K):
    st, en = start[i].min().item() - t0, end[i].max().item() - t0
    st_max = start[i].max().item() - t0
    gap = (start[i].min().item() - end[i-1].max().item()) if i else 0
    print(f'kernel {i}: first CTA start {st/1e3:7.2f} us, last CTA start {st_max/1e3:7.2f}, last CTA end {en/1e3:7.2f} us, '
          f'gap from prev end {gap/1e3:6.2f} us, span {(en-st)/1e3:6.2f} us')
