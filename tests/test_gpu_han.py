"""GPU parity of the HAN widening: residual groups in the trunk kernels + the layer-attention and
channel-spatial-attention kernels of csrc/han.cu, against outputs of the unmodified reference (tests/golden/han.npz)."""
import ctypes
import os

import numpy as np
import pytest
import torch

import recipe
from oracle import sr_torch_cpu

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
MODES = {'per-layer': (0, 0), 'dataflow': (1, 0), 'cluster': (1, 1)}


def _lib():
    from rumpy_b200 import _lib
    lib = _lib.load()
    return lib


def _han(nb, scale, sd):
    from rumpy_b200.SISR.models.advanced.architectures import HAN
    net = HAN(n_resblocks=nb, scale=scale)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return net.to(DEV).eval()


@pytest.mark.parametrize('name', list(recipe.HCASES))
def test_han_matches_reference_golden_in_every_trunk_mode(golden_dir, name):
    gold = np.load(os.path.join(golden_dir, 'han.npz'))
    nb, scale, sd, x = recipe.hcase_tensors(name)
    net = _han(nb, scale, sd)
    xt = torch.from_numpy(x).to(DEV)
    ref = gold[name + '::out']
    lib = _lib()
    outs = {}
    try:
        for mode, (trunk, cluster) in MODES.items():
            eng = net.native_engine()
            eng.set_option('trunk', trunk)
            eng.set_option('cluster', cluster)
            eng._ws.clear()
            eng._graphs.clear()
            eng._last_infer_shape = None
            with torch.no_grad():
                a = net(xt).clone()
                b = net(xt).clone()      # CUDA-graph replay
            assert bool((a == b).all()), f'{mode}: graph replay differs from the eager forward'
            outs[mode] = a.cpu().numpy()
            err = float(np.abs(outs[mode] - ref).max())
            assert err <= 1e-2, f'{name} [{mode}]: max-abs {err} vs the reference output'
    finally:
        net.native_engine().set_option('trunk', 1)
        net.native_engine().set_option('cluster', 1)


def test_han_handler_cfg2_batch_vs_oracle(tmp_path):
    """HANHandler (registry name 'han', 10 groups x 20 blocks as the reference locks it) at BASELINE configs[1]'s batch
    through run_eval, against the CPU oracle."""
    from rumpy_b200.shared_framework.models import define_model
    h = define_model('han', device=0, model_save_dir=str(tmp_path), eval_mode=True, scale=4)
    assert h.model_name == 'han'
    spec = recipe.han_spec(20)
    assert [(k, tuple(v.shape)) for k, v in h.net.state_dict().items()] == [(k, tuple(s)) for k, s in spec]
    sd = recipe.make_weights(spec, seed=85)
    sd['la.gamma'] = np.array([0.3], dtype=np.float32)
    sd['csa.gamma'] = np.array([0.5], dtype=np.float32)
    sd['csa.conv.weight'] = np.random.RandomState(86).uniform(-0.5, 0.5, (1, 1, 3, 3, 3)).astype(np.float32)
    h.net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    x = recipe.make_input((16, 3, 48, 48), 87)
    out, _, _ = h.run_eval(torch.from_numpy(x))
    assert _lib().rumpy_net_trunk_mode(h.net.native_engine().handle) == 2      # cluster kernel for the groups
    ref = sr_torch_cpu.han_forward({k: torch.from_numpy(v) for k, v in sd.items()}, torch.from_numpy(x), 10, 20, 4).numpy()
    assert float(np.abs(out.numpy() - ref).max()) <= 1e-2


@pytest.mark.parametrize('bwd', [1, 0], ids=['dataflow-backward', 'per-layer-backward'])
@pytest.mark.parametrize('name', list(recipe.HCASES))
def test_han_gradients_and_adam_steps_vs_reference_golden(golden_dir, name, bwd):
    """HAN training: trunk forward in the dataflow kernel, backward in the dataflow kernel or through the per-layer kernels,
    with the layer-attention gradients injected at the group boundaries; every gradient (convs, channel attention, csa.conv, both gammas) against the
    reference autograd (<= 5 % of the tensor's max magnitude -- the layer attention's softmax over 10^3-sized energies
    amplifies the bf16 rounding of the stacked features a little beyond the 3 % the plain trunk needs -- and cosine
    >= 0.999), then 3 Adam steps against the reference's losses (<= 1 %)."""
    from rumpy_b200 import train_native
    from rumpy_b200.optim import FusedAdam
    gold = np.load(os.path.join(golden_dir, 'han.npz'))
    nb, scale, sd, x = recipe.hcase_tensors(name)
    y = recipe.make_input((x.shape[0], 3, x.shape[2] * scale, x.shape[3] * scale), recipe.HCASES[name][4] + 1000)
    _check_han_training(gold, name, nb, scale, sd, x, y, train_native, FusedAdam, bwd)


def _check_han_training(gold, name, nb, scale, sd, x, y, train_native, FusedAdam, bwd):
    net = _han(nb, scale, sd).train()
    eng = net.native_engine()
    eng.set_option('trunk_bwd', bwd)
    xt, yt = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
    out = eng.forward(xt, training=True)
    loss, dy = train_native.l1_loss(out, yt, want_grad=True)
    grads = eng.backward(xt, dy)
    assert abs(loss.item() - float(gold[name + '::loss'])) <= 0.01 * float(gold[name + '::loss'])
    for (k, _), g in zip(net.named_parameters(), grads):
        ref = gold[name + '::gradsub::' + k]
        got = recipe.subsample(g.cpu().numpy())
        scale_ = max(float(np.abs(ref).max()), 1e-12)
        assert np.abs(got - ref).max() <= 0.05 * scale_, (k, float(np.abs(got - ref).max()), scale_)
        if np.linalg.norm(ref) < 1e-12:         # dead ReLU in a channel-attention block: the gradient is exactly zero
            assert float(np.abs(got).max()) <= 1e-9, k
        elif ref.size > 1:
            cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
            assert cos >= (0.999 if ref.size >= 64 else 0.995), (k, cos)   # 27-tap csa.conv.weight: few elements
    net = _han(nb, scale, sd).train()
    net.native_engine().set_option('trunk_bwd', bwd)
    opt = FusedAdam(list(net.parameters()), lr=1e-4)
    losses = [train_native.train_step(net, opt, xt, yt)[0].item() for _ in range(3)]
    np.testing.assert_allclose(losses, gold[name + '::train_losses'], rtol=0.01)


def test_qhan_forward_and_gradients_vs_reference_golden(golden_dir):
    """Q-HAN (reference attention_manipulators/architectures.py:643-760): Q-RCAN's residual groups inside HAN -- the
    native net is the same executor with both features switched on.  Forward in all trunk modes and every gradient
    (q-layers, attention modules, convs) against the unmodified reference."""
    from rumpy_b200 import train_native
    from rumpy_b200.SISR.models.attention_manipulators.architectures import QHAN
    gold = np.load(os.path.join(golden_dir, 'han.npz'))
    kw, has_q, sd, x, meta = recipe.qhcase_tensors()
    net = QHAN(**kw)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    net = net.to(DEV).eval()
    xt = torch.from_numpy(x).to(DEV)
    attrs = torch.from_numpy(meta).unsqueeze(2).unsqueeze(3).to(DEV)
    lib = _lib()
    try:
        for mode, (trunk, cluster) in MODES.items():
            eng = net.native_engine()
            eng.set_option('trunk', trunk)
            eng.set_option('cluster', cluster)
            eng._ws.clear(); eng._graphs.clear(); eng._last_infer_shape = None
            with torch.no_grad():
                out = net(xt, attrs).cpu().numpy()
            assert float(np.abs(out - gold['qhan::out']).max()) <= 1e-2, mode
    finally:
        net.native_engine().set_option('trunk', 1)
        net.native_engine().set_option('cluster', 1)
    net.train()
    eng = net.native_engine()
    y = recipe.make_input((x.shape[0], 3, x.shape[2] * kw['scale'], x.shape[3] * kw['scale']), recipe.QHCASE['xseed'] + 1000)
    eng.set_metadata(attrs, x.shape[0])
    o = eng.forward(xt, training=True)
    loss, dy = train_native.l1_loss(o, torch.from_numpy(y).to(DEV), want_grad=True)
    grads = eng.backward(xt, dy)
    assert abs(loss.item() - float(gold['qhan::loss'])) <= 0.01 * float(gold['qhan::loss'])
    n_q = 0
    for (k, _), g in zip(net.named_parameters(), grads):
        ref = gold['qhan::gradsub::' + k]
        got = recipe.subsample(g.cpu().numpy())
        scale_ = max(float(np.abs(ref).max()), 1e-12)
        # + a 5e-8 floor: a nearly dead channel-attention block has gradients of 1e-7, where bf16 noise dominates
        assert np.abs(got - ref).max() <= 0.05 * scale_ + 5e-8, (k, float(np.abs(got - ref).max()), scale_)
        n_q += 'q_node' in k
    assert n_q == 4 * sum(has_q)
