"""Regression tests (-m gpu) for stale-state bugs of the host layer (engine.py / optim.py / blocks_native.py /
train_native.py): the packed bf16 weights, CUDA graphs and block caches must follow every way the fp32 parameters
can change (load_state_dict, in-place edits, a foreign optimiser, the fused Adam), and a graph must never outlive
the workspace it was captured against."""
import os

import numpy as np
import pytest
import torch

import recipe

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _rcan(seed, **kw):
    from rumpy_b200.SISR.models.advanced.architectures import RCAN
    net = RCAN(n_resgroups=kw.get('g', 1), n_resblocks=kw.get('b', 2))
    spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    sd = {k: torch.from_numpy(v) for k, v in recipe.make_weights(spec, seed=seed).items()}
    net.load_state_dict(sd, strict=True)
    return net.to(DEV).eval(), sd


def _x(shape, seed=8):
    return torch.from_numpy(recipe.make_input(shape, seed=seed)).to(DEV)


def test_load_state_dict_after_first_forward_repacks_weights():
    net, _ = _rcan(8)
    fresh, sd2 = _rcan(9)
    x = _x((2, 3, 16, 16))
    with torch.no_grad():
        a0 = net(x).clone()
        a1 = net(x).clone()                      # second call: CUDA-graph replay
        net.load_state_dict(sd2, strict=True)    # writes through the parameters (their own version counters)
        b = net(x).clone()
        b2 = net(x).clone()
        want = fresh(x).clone()
    assert torch.equal(a0, a1)
    assert not torch.equal(a0, want)
    assert torch.equal(b, want) and torch.equal(b2, want), 'stale packed weights after load_state_dict'


def test_in_place_parameter_edit_is_seen():
    net, _ = _rcan(8)
    x = _x((2, 3, 16, 16))
    with torch.no_grad():
        a = net(x).clone()
        net(x)
        net.tail[1].weight.mul_(0.5)
        net.tail[1].bias.mul_(0.5)
        b = net(x).clone()
    assert torch.allclose(b, 0.5 * a, rtol=2e-2, atol=1e-3) and not torch.equal(a, b)


def test_graph_is_dropped_when_its_workspace_is_evicted():
    """Shapes A, A (graph captured), B (A's workspace released), A, A: the last call must not replay a graph that
    was captured against the released workspace."""
    net, _ = _rcan(8)
    xa, xb = _x((2, 3, 16, 16)), _x((1, 3, 24, 40), seed=9)
    eng = net.native_engine()
    with torch.no_grad():
        a0 = net(xa).clone()
        a1 = net(xa).clone()
        assert eng._graphs, 'second call with one shape should have captured a graph'
        net(xb)
        assert not eng._graphs, 'graphs must go with the workspace they were captured against'
        junk = torch.full((64 << 20,), 7, dtype=torch.uint8, device=DEV)     # reuse the freed block
        a2 = net(xa).clone()
        a3 = net(xa).clone()
        del junk
    assert torch.equal(a0, a1) and torch.equal(a0, a2) and torch.equal(a0, a3)


def test_foreign_optimizer_step_reaches_the_packed_weights_and_is_averaged_and_clipped():
    """torch's RMSprop (the reference's optimizer_type switch, base_architecture.py:79-91) on the native gradients:
    the clip coefficient is applied to what the optimiser sees, and the next forward uses the stepped weights."""
    from rumpy_b200 import train_native
    net, _ = _rcan(8)
    net.train()
    x, y = _x((2, 3, 16, 16)), _x((2, 3, 64, 64), seed=10)
    opt = torch.optim.RMSprop(list(net.parameters()), lr=1e-3)
    eng = net.native_engine()
    out = eng.forward(x, training=True)
    _, dy = train_native.l1_loss(out, y, want_grad=True)
    eng.backward(x, dy)
    g0 = eng.flat_grads.clone()
    clip = 0.25 * float(g0.norm())
    before = [p.detach().clone() for p in net.parameters()]
    train_native.train_step(net, opt, x, y, grad_clip=clip)
    seen = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert abs(float(seen.norm()) - clip) <= 1e-3 * clip, 'RMSprop must see the clipped gradient'
    assert any(not torch.equal(a, p.detach()) for a, p in zip(before, net.parameters()))
    # the stepped weights are what the next forward computes with
    from rumpy_b200.SISR.models.advanced.architectures import RCAN
    fresh = RCAN(n_resgroups=1, n_resblocks=2)
    fresh.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()}, strict=True)
    fresh = fresh.to(DEV).eval()
    net.eval()
    with torch.no_grad():
        assert torch.equal(net(x), fresh(x)), 'stale packed weights after a foreign optimiser step'


def test_fused_adam_zero_grad_clears_and_block_caches_follow_the_fused_step():
    from rumpy_b200 import train_native
    from rumpy_b200.optim import FusedAdam
    net, _ = _rcan(8)
    x, y = _x((2, 3, 16, 16)), _x((2, 3, 64, 64), seed=10)
    xf = torch.rand((2, 64, 16, 16), device=DEV)
    with torch.no_grad():
        blk0 = net.body[0].body[0](xf).clone()          # stand-alone RCAB forward: packs and caches its weights
    net.train()
    opt = FusedAdam(list(net.parameters()), lr=1e-2)
    for _ in range(3):
        train_native.train_step(net, opt, x, y)
    assert float(opt.flat_g.abs().sum()) > 0
    opt.zero_grad()
    assert float(opt.flat_g.abs().sum()) == 0.0 and float(net.head[0].weight.grad.abs().sum()) == 0.0
    net.eval()
    from rumpy_b200.SISR.models.advanced.architectures import RCAN
    fresh = RCAN(n_resgroups=1, n_resblocks=2)
    fresh.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()}, strict=True)
    fresh = fresh.to(DEV).eval()
    with torch.no_grad():
        blk1 = net.body[0].body[0](xf)
        want = fresh.body[0].body[0](xf)
    assert not torch.equal(blk0, blk1)
    assert torch.equal(blk1, want), 'stand-alone block used packed weights from before the fused Adam steps'


def test_qmodel_checkpoint_is_written_without_metadata_keys(tmp_path):
    """Reference attention_manipulators/__init__.py:167-175 saves unconditionally; `extra_channels` passed directly
    never sets `metadata_keys_used_in_training`."""
    from rumpy_b200.shared_framework.models import define_model
    h = define_model('qrcan', device=0, model_save_dir=str(tmp_path), eval_mode=False, scale=4, n_resgroups=1,
                     n_resblocks=2, metadata=['blur_kernel'], style='standard', include_q_layer=True)
    h.save_model('train_model')
    path = os.path.join(str(tmp_path), 'train_model_0')
    assert os.path.exists(path)
    state = torch.load(path, map_location='cpu', weights_only=False)
    assert state['model_name'] == 'qrcan' and 'metadata_keys_used_in_training' not in state


@pytest.mark.parametrize('shape,mode', [((1, 3, 100, 200), 1), ((2, 3, 48, 48), 2), ((1, 3, 300, 520), 0)])
def test_frames_in_flight_gives_the_sequential_results(shape, mode):
    """parallel.FramesInFlight (two / three frames in flight on as many streams and engines sharing the parameters) must
    return exactly what one-frame-at-a-time inference returns, in order -- for every trunk path (dataflow kernel,
    cluster kernel, one kernel per layer): the kernels of different frames run concurrently on one GPU."""
    from rumpy_b200 import parallel
    net, _ = _rcan(8, g=2, b=2)
    frames = [_x(shape, seed=20 + s) for s in range(5)]
    with torch.no_grad():
        net.native_engine().forward(frames[0])
    assert net.native_engine().lib.rumpy_net_trunk_mode(net.native_engine().handle) == mode
    with torch.no_grad():
        want = [net.native_engine().forward(f).clone() for f in frames]
    got = parallel.FramesInFlight(net, depth=2).run(frames)
    torch.cuda.synchronize()
    assert len(got) == 5 and all(torch.equal(a, b) for a, b in zip(want, got))
    seen = {}
    parallel.FramesInFlight(net, depth=3).run(frames, consume=lambda i, out: seen.__setitem__(i, out.clone()))
    torch.cuda.synchronize()
    assert sorted(seen) == [0, 1, 2, 3, 4] and all(torch.equal(want[i], seen[i]) for i in range(5))
