"""Under torchrun: RCAN train-step time (device events, max over ranks) with the gradient all-reduce off / after the
backward / overlapped with 4 or 8 weight-gradient chunks -- separates straggler GPUs from exposed communication."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import torch.distributed as dist
import recipe
from rumpy_b200 import _lib, parallel, train_native
from rumpy_b200.optim import FusedAdam
from rumpy_b200.SISR.models.advanced.architectures import RCAN

local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
lib = _lib.load()
x, y = torch.rand((16, 3, 64, 64), device=dev), torch.rand((16, 3, 256, 256), device=dev)
sd = {k: torch.from_numpy(v) for k, v in recipe.make_weights(recipe.rcan_spec(), seed=8).items()}


class Plain(parallel.GradAllReduce):          # all-reduce after the whole backward (no overlap)
    def chunked(self, flat, chunks):
        return self(flat)


for name, chunks, ar in (('no all-reduce', 4, None), ('after backward', 4, Plain()), ('overlapped, 4 chunks', 4,
                         parallel.GradAllReduce()), ('overlapped, 8 chunks', 8, parallel.GradAllReduce()),
                         ('overlapped, 4 chunks, 4 MB buckets', 4, parallel.GradAllReduce(bucket_bytes=4 << 20))):
    net = RCAN()
    net.load_state_dict(sd)
    net = net.to(dev).train()
    net.native_engine().set_option('wgrad_chunks', chunks)
    opt = FusedAdam(list(net.parameters()), lr=1e-4)
    for _ in range(5):
        train_native.train_step(net, opt, x, y, allreduce=ar)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); train_native.train_step(net, opt, x, y, allreduce=ar); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    t = torch.tensor([ms, -ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'{world} GPUs, {name}: slowest rank {t[0].item():.3f} ms/step, fastest {-t[1].item():.3f}', flush=True)
    del net, opt
    torch.cuda.empty_cache()
dist.destroy_process_group()
