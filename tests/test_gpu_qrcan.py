"""GPU parity of the meta-attention widening (Q-RCAN, Q-EDSR): the native QRCAN / QEDSR modules and their handlers
against outputs of the unmodified reference (tests/golden/qrcan.npz) and the CPU oracle, in all three trunk modes."""
import ctypes
import os

import numpy as np
import pytest
import torch

import recipe
from oracle import sr_torch_cpu

pytestmark = pytest.mark.gpu

MODES = {'per-layer': (0, 0), 'dataflow': (1, 0), 'cluster': (1, 1)}


def _dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda:0')


def _lib():
    from rumpy_b200 import _lib
    lib = _lib.load()
    return lib


def _qrcan(kw, sd, cls='QRCAN'):
    from rumpy_b200.SISR.models.attention_manipulators import architectures
    net = getattr(architectures, cls)(**kw)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return net.to(_dev()).eval()


@pytest.mark.parametrize('name', list(recipe.QCASES) + list(recipe.QECASES))
def test_qrcan_matches_reference_golden_in_every_trunk_mode(golden_dir, name):
    gold = np.load(os.path.join(golden_dir, 'qrcan.npz'))
    if name in recipe.QCASES:
        kw, has_q, sd, x, meta = recipe.qcase_tensors(name)
        net = _qrcan(kw, sd)
    else:
        kw, has_q, sd, x, meta = recipe.qecase_tensors(name)
        net = _qrcan(kw, sd, 'QEDSR')
    attrs = torch.from_numpy(gold[name + '::attributes']).to(_dev())
    xt = torch.from_numpy(x).to(_dev())
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
                a = net(xt, attrs).clone()       # eager launch
                b = net(xt, attrs).clone()       # CUDA-graph replay of the same shape
            assert bool((a == b).all()), f'{mode}: graph replay differs from the eager forward'
            outs[mode] = a.cpu().numpy()
            err = float(np.abs(outs[mode] - ref).max())
            assert err <= 1e-2, f'{name} [{mode}]: max-abs {err} vs the reference output'
    finally:
        net.native_engine().set_option('trunk', 1)
        net.native_engine().set_option('cluster', 1)
    for mode in ('dataflow', 'cluster'):
        assert float(np.abs(outs[mode] - outs['per-layer']).max()) <= 5e-3


def test_qrcan_metadata_changes_output_through_graph_replay():
    """New metadata with the same batch shape must reach the replayed CUDA graph (static metadata buffer)."""
    kw, has_q, sd, x, meta = recipe.qcase_tensors('qrcan_blur_q')
    net = _qrcan(kw, sd)
    xt = torch.from_numpy(x).to(_dev())
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    for seed in (1, 2, 3):
        m = recipe.make_input(meta.shape, seed)
        attrs = torch.from_numpy(m).unsqueeze(2).unsqueeze(3)
        with torch.no_grad():
            out = net(xt, attrs.to(_dev())).cpu().numpy()
        ref = sr_torch_cpu.qrcan_forward(tsd, torch.from_numpy(x), attrs, kw['n_resgroups'], kw['n_resblocks'],
                                         kw['scale'], kw['style']).numpy()
        assert float(np.abs(out - ref).max()) <= 1e-2, f'metadata seed {seed}'


def test_qrcan_handler_run_eval_with_metadata_keys(tmp_path):
    """QRCANHandler.run_eval(x, metadata=..., metadata_keys=...) selects the handler's metadata columns, builds
    the [N,M,1,1] vector (reference attention_manipulators/__init__.py:87-108) and runs the native trunk."""
    from rumpy_b200.shared_framework.models import define_model
    h = define_model('qrcan', device=0, model_save_dir=str(tmp_path), eval_mode=True, scale=4, style='standard',
                     metadata=['blur_sigma', 'noise'], include_q_layer=True, n_resgroups=1, n_resblocks=2)
    assert h.model_name == 'qrcan' and h.num_metadata == 2 and h.colorspace == 'augmented_rgb'
    spec = [(k, tuple(v.shape)) for k, v in h.net.state_dict().items()]
    sd = recipe.make_weights(spec, seed=70)
    h.net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    x = recipe.make_input((2, 3, 16, 12), 71)
    table = torch.from_numpy(recipe.make_input((2, 3), 72))      # columns: blur_sigma, jpeg (unused), noise
    keys = [('blur_sigma',), ('jpeg_quality',), ('noise',)]
    out, loss, _ = h.run_eval(torch.from_numpy(x), metadata=table, metadata_keys=keys)
    assert tuple(out.shape) == (2, 3, 64, 48) and out.device.type == 'cpu' and loss is None
    attrs = table[:, [0, 2]].unsqueeze(2).unsqueeze(3)
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    ref = sr_torch_cpu.qrcan_forward(tsd, torch.from_numpy(x), attrs, 1, 2, 4, 'standard').numpy()
    assert float(np.abs(out.numpy() - ref).max()) <= 1e-2
    assert h.metadata_keys_used_in_training == ['blur_sigma', 'jpeg_quality', 'noise']


def test_qrcan_full_size_uses_cluster_kernel():
    """Sample q-rcan.toml configuration (10 groups x 20 blocks, blur-kernel metadata M=10, q-node in every RCAB)
    at BASELINE configs[1]'s batch: same launch structure as RCAN plus one metadata kernel; oracle parity."""
    from rumpy_b200.SISR.models.attention_manipulators.architectures import QRCAN
    net = QRCAN(style='standard', num_metadata=10, include_q_layer=True)
    spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    sd = recipe.make_weights(spec, seed=80)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    net = net.to(_dev()).eval()
    x = recipe.make_input((16, 3, 48, 48), 81)
    attrs = torch.from_numpy(recipe.make_input((16, 10), 82)).unsqueeze(2).unsqueeze(3)
    with torch.no_grad():
        out = net(torch.from_numpy(x).to(_dev()), attrs.to(_dev())).cpu().numpy()
    eng = net.native_engine()
    assert _lib().rumpy_net_trunk_mode(eng.handle) == 2
    assert _lib().rumpy_net_num_launches(eng.handle) == 6      # head, metadata, trunk, 2 upsampler convs, tail
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    ref = sr_torch_cpu.qrcan_forward(tsd, torch.from_numpy(x), attrs, 10, 20, 4, 'standard').numpy()
    assert float(np.abs(out - ref).max()) <= 1e-2


def test_qedsr_handler_full_width_per_layer_path(tmp_path):
    """Sample q-edsr.toml configuration (256 features, blur-kernel metadata M=10, ReLU q-layers) with 2 blocks:
    the 256-channel trunk runs on the per-layer kernels, the multiplier rides in the conv epilogue."""
    from rumpy_b200.shared_framework.models import define_model
    h = define_model('qedsr', device=0, model_save_dir=str(tmp_path), eval_mode=True, scale=4, num_blocks=2,
                     num_features=256, res_scale=0.1, metadata=['blur_kernel'], q_layer_nonlinearity=True)
    assert h.model_name == 'qedsr' and h.num_metadata == 10
    spec = [(k, tuple(v.shape)) for k, v in h.net.state_dict().items()]
    assert spec[4][0] == 'body.0.body.0.weight' and spec[8] == ('body.0.attention_layer.attribute_integrator.0.weight',
                                                                (128, 10, 1, 1))
    sd = recipe.make_weights(spec, seed=90)
    h.net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    x = recipe.make_input((2, 3, 24, 20), 91)
    attrs = torch.from_numpy(recipe.make_input((2, 10), 92)).unsqueeze(2).unsqueeze(3)
    out, _, _ = h.run_eval(torch.from_numpy(x), extra_channels=attrs)
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    ref = sr_torch_cpu.qedsr_forward(tsd, torch.from_numpy(x), attrs, 2, 0.1, 4).numpy()
    assert float(np.abs(out.numpy() - ref).max()) <= 1e-2


# ----------------------------------------------------------------------------- training (Q-RCAN style 'standard')
def _ytrain(name, kw, x):
    return recipe.make_input((x.shape[0], 3, x.shape[2] * kw['scale'], x.shape[3] * kw['scale']),
                             recipe.QCASES[name][3] + 1000)


@pytest.mark.parametrize('trunk', [1, 0], ids=['dataflow-kernels', 'per-layer-kernels'])
@pytest.mark.parametrize('name', ['qrcan_blur_q', 'qrcan_selective', 'qrcan_wide_meta', 'qrcan_modulate'])
def test_qrcan_gradients_vs_reference_golden(golden_dir, name, trunk):
    """Every parameter gradient (convs, channel attention, q-layers) of the native backward against the reference's
    autograd: <= 3 % of the tensor's max magnitude and cosine >= 0.999 (bf16 operands, fp32 accumulate).  Both
    backward implementations: the dataflow kernels and the per-layer kernels (shapes beyond 4 tiles per SM)."""
    from rumpy_b200 import train_native
    gold = np.load(os.path.join(golden_dir, 'qrcan.npz'))
    kw, has_q, sd, x, meta = recipe.qcase_tensors(name)
    net = _qrcan(kw, sd).train()
    net.native_engine().set_option('trunk', trunk)
    _check_qrcan_gradients(net, gold, name, kw, has_q, x, train_native)


def _check_qrcan_gradients(net, gold, name, kw, has_q, x, train_native):
    eng = net.native_engine()
    xt = torch.from_numpy(x).to(_dev())
    yt = torch.from_numpy(_ytrain(name, kw, x)).to(_dev())
    eng.set_metadata(torch.from_numpy(gold[name + '::attributes']), x.shape[0])
    out = eng.forward(xt, training=True)
    loss, dy = train_native.l1_loss(out, yt, want_grad=True)
    grads = eng.backward(xt, dy)
    assert abs(loss.item() - float(gold[name + '::loss'])) <= 0.01 * float(gold[name + '::loss'])
    n_q = 0
    for (k, _), g in zip(net.named_parameters(), grads):
        ref = gold[name + '::gradsub::' + k]
        got = recipe.subsample(g.cpu().numpy())
        scale = max(float(np.abs(ref).max()), 1e-12)
        assert np.abs(got - ref).max() <= 0.03 * scale, k
        cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
        assert cos >= 0.999, (k, cos)
        n_q += 'q_node' in k
    assert n_q == 4 * sum(has_q)


def test_qrcan_three_adam_steps_and_autograd_path(golden_dir):
    from rumpy_b200 import train_native
    from rumpy_b200.optim import FusedAdam
    name = 'qrcan_blur_q'
    gold = np.load(os.path.join(golden_dir, 'qrcan.npz'))
    kw, has_q, sd, x, meta = recipe.qcase_tensors(name)
    attrs = torch.from_numpy(gold[name + '::attributes']).to(_dev())
    xt = torch.from_numpy(x).to(_dev())
    yt = torch.from_numpy(_ytrain(name, kw, x)).to(_dev())
    # autograd.Function path: net(x, metadata) under grad + a torch-side loss
    net = _qrcan(kw, sd).train()
    out = net(xt, attrs)
    assert out.requires_grad
    (out - yt).abs().mean().backward()
    for k, p in net.named_parameters():
        ref = gold[name + '::gradsub::' + k]
        assert np.abs(recipe.subsample(p.grad.cpu().numpy()) - ref).max() <= 0.03 * max(float(np.abs(ref).max()), 1e-12), k
    # native train step x3 against the reference's Adam losses
    net = _qrcan(kw, sd).train()
    opt = FusedAdam(list(net.parameters()), lr=1e-4)
    losses = [train_native.train_step(net, opt, xt, yt, metadata=attrs)[0].item() for _ in range(3)]
    np.testing.assert_allclose(losses, gold[name + '::train_losses'], rtol=0.01)
    net.eval()
    with torch.no_grad():
        o = net(xt, attrs).cpu().numpy()
    assert np.abs(o - gold[name + '::out_after3']).max() <= 1e-2


def test_qrcan_handler_run_train_loss_curve_vs_oracle(tmp_path):
    """QRCANHandler.run_train(x, y, metadata, metadata_keys) for 60 steps against the CPU oracle's Adam on the same
    batches: per-step loss within 1 %."""
    from rumpy_b200.shared_framework.models import define_model
    h = define_model('qrcan', device=0, model_save_dir=str(tmp_path), eval_mode=False, scale=4, style='standard',
                     metadata=['blur_kernel'], include_q_layer=True, n_resgroups=1, n_resblocks=2, lr=1e-4)
    spec = [(k, tuple(v.shape)) for k, v in h.net.state_dict().items()]
    sd = recipe.make_weights(spec, seed=95)
    h.net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    keys = [('blur_kernel',)]
    batches = [(recipe.make_input((2, 3, 16, 16), 300 + i), recipe.make_input((2, 3, 64, 64), 400 + i),
                recipe.make_input((2, 10), 500 + i)) for i in range(4)]
    tr = None
    for step in range(60):
        x, y, m = batches[step % 4]
        attrs = torch.from_numpy(m).unsqueeze(2).unsqueeze(3)
        if tr is None:
            tr = sr_torch_cpu.Trainer({k: torch.from_numpy(v) for k, v in sd.items()}, 'qrcan', lr=1e-4,
                                      attributes=attrs, n_resgroups=1, n_resblocks=2, scale=4)
        tr.kw['attributes'] = attrs
        l_cpu, _ = tr.step(torch.from_numpy(x), torch.from_numpy(y))
        l_gpu, out = h.run_train(torch.from_numpy(x), torch.from_numpy(y), metadata=torch.from_numpy(m),
                                 metadata_keys=keys)
        assert abs(float(l_gpu) - l_cpu) <= 0.01 * l_cpu, (step, float(l_gpu), l_cpu)
    assert tuple(out.shape) == (2, 3, 64, 64)


def test_modulate_with_q_layers_training_is_rejected():
    """attributes * sigmoid(q-layer) is inference only (the reference's handler cannot build that combination)."""
    from rumpy_b200 import _lib
    from rumpy_b200.SISR.models.attention_manipulators.architectures import QRCAN
    net = QRCAN(n_resgroups=1, n_resblocks=1, style='modulate', num_metadata=64, include_q_layer=True).to(_dev()).train()
    with pytest.raises(_lib.RumpyB200Error, match='inference only'):
        net(torch.rand(1, 3, 16, 16, device=_dev()), torch.rand(1, 64, 1, 1, device=_dev()))


@pytest.mark.parametrize('name', list(recipe.QECASES))
def test_qedsr_gradients_and_adam_steps_vs_reference_golden(golden_dir, name):
    """Q-EDSR training (64-channel trunk through the dataflow kernel, 256-channel through the per-layer kernels): every
    gradient incl. the q-layers vs the reference autograd, then 3 Adam steps vs the reference's losses."""
    from rumpy_b200 import train_native
    from rumpy_b200.optim import FusedAdam
    gold = np.load(os.path.join(golden_dir, 'qrcan.npz'))
    kw, has_q, sd, x, meta = recipe.qecase_tensors(name)
    net = _qrcan(kw, sd, 'QEDSR').train()
    eng = net.native_engine()
    attrs = torch.from_numpy(gold[name + '::attributes']).to(_dev())
    xt = torch.from_numpy(x).to(_dev())
    y = recipe.make_input((x.shape[0], 3, x.shape[2] * kw['scale'], x.shape[3] * kw['scale']),
                          recipe.QECASES[name][3] + 1000)
    yt = torch.from_numpy(y).to(_dev())
    eng.set_metadata(attrs, x.shape[0])
    out = eng.forward(xt, training=True)
    loss, dy = train_native.l1_loss(out, yt, want_grad=True)
    grads = eng.backward(xt, dy)
    assert abs(loss.item() - float(gold[name + '::loss'])) <= 0.01 * float(gold[name + '::loss'])
    n_q = 0
    for (k, _), g in zip(net.named_parameters(), grads):
        ref = gold[name + '::gradsub::' + k]
        got = recipe.subsample(g.cpu().numpy())
        scale = max(float(np.abs(ref).max()), 1e-12)
        assert np.abs(got - ref).max() <= 0.03 * scale, (k, float(np.abs(got - ref).max()), scale)
        cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
        assert cos >= 0.999, (k, cos)
        n_q += 'attention_layer' in k
    assert n_q == 4 * sum(has_q)
    net = _qrcan(kw, sd, 'QEDSR').train()
    opt = FusedAdam(list(net.parameters()), lr=1e-4)
    losses = [train_native.train_step(net, opt, xt, yt, metadata=attrs)[0].item() for _ in range(3)]
    np.testing.assert_allclose(losses, gold[name + '::train_losses'], rtol=0.01)
    net.eval()
    with torch.no_grad():
        o = net(xt, attrs).cpu().numpy()
    # 1e-2 is the forward tolerance; three sign-like Adam steps on the 256-channel net (2304-term bf16 dot products,
    # random init) add a little drift on top of its ~9e-3 forward error
    assert np.abs(o - gold[name + '::out_after3']).max() <= (1.5e-2 if kw['num_features'] > 64 else 1e-2)
