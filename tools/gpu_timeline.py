"""Per-CTA clock64 timeline of the conv kernel (debug hook rumpy_debug_set_timeline)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rumpy_b200 import ops, _lib

names = {0: 'entry', 1: 'setup_sync_done', 2: 'B_loads_issued', 3: 'B_landed(mma)', 4: 'first_A_landed(mma)',
         5: 'tile0_mma_issued+commit', 6: 'epi_tile0_tmem_full', 8: 'epi_tile0_staged', 9: 'epi_tile1_tmem_full',
         10: 'epi_tile1_staged', 11: 'epi_done_before_store_wait', 12: 'store_wait_done', 13: 'final_sync_done'}
lib = _lib.load()
lib.rumpy_debug_set_timeline.argtypes = [ctypes.c_void_p]
dev = torch.device('cuda:0')
for mode in ('relu_bf16', 'pool_f32'):
    N, H, W, C = 16, 48, 48, 64
    x = torch.rand((N, H, W, C), device=dev).to(torch.bfloat16)
    w = (torch.rand((C, C, 3, 3), device=dev) - 0.5) / 24
    b = torch.rand((C,), device=dev)
    wp = ops.pack_conv3x3(w)
    y = torch.empty_like(x)
    yf = torch.empty((N, H, W, C), device=dev)
    pp = torch.empty((N * ops.tiles_per_image(H, W), 2, C), device=dev)
    dbg = torch.zeros((148, 16), dtype=torch.int64, device=dev)
    def run():
        if mode == 'relu_bf16':
            ops.conv3x3(x, wp, b, out_bf16=y, N=N, H=H, W=W, Cin=C, Cout=C, relu=True)
        else:
            ops.conv3x3(x, wp, b, out_f32=yf, pool_partial=pp, N=N, H=H, W=W, Cin=C, Cout=C)
    for _ in range(3): run()
    torch.cuda.synchronize()
    lib.rumpy_debug_set_timeline(dbg.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.rumpy_debug_set_timeline(None)
    d = dbg.cpu()
    print(f'--- mode {mode}: cycles since CTA entry (CTAs 0, 1, 74, 147) and median over CTAs')
    rel = d - d[:, :1]
    for slot in sorted(names):
        col = rel[:, slot]
        med = col[col > 0].median().item() if (col > 0).any() else 0
        print(f'  {names[slot]:28s} ' + ' '.join(f'{rel[c, slot].item():8d}' for c in (0, 1, 74, 147)) + f'   median {med:8.0f}')
