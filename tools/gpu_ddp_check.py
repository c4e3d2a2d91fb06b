"""Under torchrun (>= 2 ranks, one GPU each): numerical check of the data-parallel train step that replaces
nn.DataParallel (reference base_architecture.py:70-77).

  1. the all-reduced flat gradient (chunked, overlapped with the backward, 1/world folded into Adam) times 1/world
     equals the MEAN of the ranks' own gradients (gathered separately) -- to fp32 summation order;
  2. it equals, to bf16-operand tolerance, the gradient ONE GPU computes on the concatenated global batch (what
     DataParallel's scatter / gather produces: L1 mean over equal shards = mean of the shard means);
  3. after the fused Adam step every rank holds bit-identical parameters, and they equal the single-GPU step on the
     concatenated batch to the same tolerance;
  4. grad-norm clipping acts on the averaged gradient.

Prints one JSON line on rank 0 and exits non-zero on failure:
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_ddp_check.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import recipe  # noqa: E402
from rumpy_b200 import parallel, train_native  # noqa: E402
from rumpy_b200.optim import FusedAdam  # noqa: E402
from rumpy_b200.SISR.models.advanced.architectures import EDSR, RCAN  # noqa: E402


def build(kind, dev, **kw):
    net = RCAN(**kw) if kind == 'rcan' else EDSR(**kw)
    spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    net.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(spec, seed=8).items()}, strict=True)
    return net.to(dev).train()


def local_grads(net, x, y):
    eng = net.native_engine()
    out = eng.forward(x, training=True)
    _, dy = train_native.l1_loss(out, y, want_grad=True)
    eng.backward(x, dy)
    return eng.flat_grads.clone()


def main():
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    res, ok = {}, True
    cases = [('rcan_2x3', 'rcan', dict(n_resgroups=2, n_resblocks=3), (4, 3, 32, 32)),
             ('edsr_4', 'edsr', dict(num_blocks=4), (4, 3, 24, 40))]
    if os.environ.get('FULL', '1') == '1':
        cases.append(('rcan_10x20_cfg3', 'rcan', {}, (16, 3, 64, 64)))
    for name, kind, kw, shape in cases:
        n, _, h, w = shape
        xs = [torch.from_numpy(recipe.make_input(shape, seed=100 + r)).to(dev) for r in range(world)]
        ys = [torch.from_numpy(recipe.make_input((n, 3, 4 * h, 4 * w), seed=200 + r)).to(dev) for r in range(world)]
        # (a) every rank's own gradient, gathered
        net = build(kind, dev, **kw)
        g_own = local_grads(net, xs[rank], ys[rank])
        gathered = [torch.empty_like(g_own) for _ in range(world)]
        dist.all_gather(gathered, g_own)
        g_mean = torch.stack(gathered).double().mean(0).float()
        # (b) the product path: chunked all-reduce overlapped with the backward
        ar = parallel.GradAllReduce()
        eng = net.native_engine()
        out = eng.forward(xs[rank], training=True)
        _, dy = train_native.l1_loss(out, ys[rank], want_grad=True)
        chunks = eng.backward_chunks()
        eng.backward(xs[rank], dy)
        ar.chunked(eng.flat_grads, chunks)
        torch.cuda.synchronize()
        g_dp = eng.flat_grads / world
        scale = float(g_mean.abs().max())
        e_mean = float((g_dp - g_mean).abs().max()) / scale
        # (c) one GPU, concatenated global batch
        netc = build(kind, dev, **kw)
        g_cat = local_grads(netc, torch.cat(xs), torch.cat(ys))
        e_cat = float((g_dp - g_cat).abs().max()) / scale
        cos = float((g_dp.double() * g_cat.double()).sum() / (g_dp.double().norm() * g_cat.double().norm()))
        # (d) a full train step through the product path vs the single-GPU step on the concatenated batch
        clip = 0.5 * float(g_mean.norm())
        net2, net3 = build(kind, dev, **kw), build(kind, dev, **kw)
        opt2, opt3 = FusedAdam(list(net2.parameters()), lr=1e-3), FusedAdam(list(net3.parameters()), lr=1e-3)
        p0 = opt2.flat_p.clone()
        train_native.train_step(net2, opt2, xs[rank], ys[rank], grad_clip=clip, allreduce=parallel.GradAllReduce())
        train_native.train_step(net3, opt3, torch.cat(xs), torch.cat(ys), grad_clip=clip)
        torch.cuda.synchronize()
        ps = [torch.empty_like(opt2.flat_p) for _ in range(world)]
        dist.all_gather(ps, opt2.flat_p)
        replicas_equal = all(bool(torch.equal(ps[0], p)) for p in ps[1:])
        # Adam's first step moves every weight by lr * g/(|g| + eps): compare the update directions where the
        # gradient is well above eps (tiny gradients flip sign under bf16 noise)
        d2, d3 = opt2.flat_p - p0, opt3.flat_p - p0
        big = g_cat.abs() > 1e-3 * scale
        agree = float((torch.sign(d2[big]) == torch.sign(d3[big])).float().mean())
        r = dict(rel_err_vs_mean_of_rank_grads=e_mean, rel_err_vs_single_gpu_concat=e_cat, cos_vs_single_gpu_concat=cos,
                 replicas_bit_identical_after_step=replicas_equal, update_sign_agreement_vs_single_gpu=agree)
        good = e_mean <= 1e-5 and e_cat <= 3e-2 and cos >= 0.999 and replicas_equal and agree >= 0.995
        r['ok'] = good
        ok = ok and good
        res[name] = r
        del net, netc, net2, net3, opt2, opt3
        torch.cuda.empty_cache()
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({'world': world, 'ok': bool(flag.item()), 'cases': res}))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == '__main__':
    main()
