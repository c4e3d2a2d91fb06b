"""Bicubic baseline kernel (csrc/glue.cu) on a 1080p frame (x4 -> 4320 x 7680) and on BASELINE configs[1]'s LR batch:
CUDA-event timing (printed as JSON) and, under `ncu -k regex:bicubic -c 2`, the kernel's full metric set."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rumpy_b200.shared_framework.data import bicubic_upsample_device

dev = torch.device('cuda:0')
res = {}
for name, shape, iters in [('frame_1080p_x4', (1, 3, 1080, 1920), 20), ('cfg2_batch_16x48x48_x4', (16, 3, 48, 48), 50)]:
    x = torch.rand(shape, device=dev)
    for _ in range(5):
        bicubic_upsample_device(x, 4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        bicubic_upsample_device(x, 4)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = x.numel() * 4 * 17
    res[name] = {'ms': ms, 'algorithmic_bytes': nbytes, 'gb_per_s': nbytes / ms * 1e-6}
print(json.dumps(res))
