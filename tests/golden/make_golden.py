"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference is pure Python on top of torch; it imports with two shims:
`collections.Callable` (removed in py3.10) and stub modules for third-party packages the
EDSR/RCAN trunk never uses.  Nothing from the reference is copied: it is imported, executed on
recipe inputs (tests/golden/recipe.py) and only its OUTPUTS are stored.
"""
from __future__ import annotations

import collections
import collections.abc
import importlib.abc
import importlib.machinery
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import recipe  # noqa: E402

REF = '/root/reference'


# ---------------------------------------------------------------------------- import shims
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402  (collections.Callable + stub modules for absent third-party packages)


def import_reference():
    ref_import.import_reference(REF)           # goldens are generated from the source tree in the build container
    from rumpy.SISR.models.advanced import architectures, common  # noqa
    return architectures, common


# ---------------------------------------------------------------------------- helpers
def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def checksum(a: np.ndarray):
    a64 = a.astype(np.float64)
    return np.array([a64.sum(), np.abs(a64).sum(), (a64 * a64).sum()])


def build_ref_net(arch_mod, arch, kw):
    if arch == 'rcan':
        return arch_mod.RCAN(n_resblocks=kw['n_resblocks'], n_resgroups=kw['n_resgroups'], n_feats=kw['n_feats'],
                             scale=kw['scale'])
    return arch_mod.EDSR(net_features=kw['n_feats'], num_blocks=kw['num_blocks'], scale=kw['scale'],
                         res_scale=kw['res_scale'])


def net_case(arch_mod, name):
    arch, kw, sd, x, y = recipe.case_tensors(name)
    torch.manual_seed(0)
    net = build_ref_net(arch_mod, arch, kw)
    # strict load: proves recipe key names / shapes / ORDER equal the reference's
    assert list(net.state_dict().keys()) == list(sd.keys()), 'key order mismatch vs reference'
    net.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
    net.train()
    xt, yt = t(x), t(y)
    out = net(xt)
    loss = torch.nn.L1Loss()(out, yt)
    loss.backward()
    rec = {'out': out.detach().numpy(), 'loss': np.float32(loss.item())}
    grads = {k: p.grad.numpy() for k, p in net.named_parameters()}
    keys = list(grads.keys())
    rec['grad_keys'] = np.array(keys)
    rec['grad_checksums'] = np.stack([checksum(grads[k]) for k in keys])
    # every gradient tensor, strided-subsampled to <= ~4k values (recipe.subsample)
    for k in keys:
        rec['gradsub::' + k] = recipe.subsample(grads[k]).copy()

    # 3 Adam steps through the reference's own optimiser recipe (base_architecture.py:93-95, 425-440)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, net.parameters()), lr=1e-4)
    losses = []
    for _ in range(3):
        o = net(xt)
        l = torch.nn.L1Loss()(o, yt)
        opt.zero_grad()
        l.backward()
        opt.step()
        losses.append(l.item())
    rec['train_losses'] = np.array(losses, dtype=np.float32)
    sd_after = net.state_dict()
    rec['after3::' + keys[0]] = sd_after[keys[0]].numpy()
    rec['after3::' + keys[-1]] = sd_after[keys[-1]].numpy()
    rec['after3sub::' + keys[2]] = recipe.subsample(sd_after[keys[2]].numpy()).copy()
    rec['after3_checksums'] = np.stack([checksum(sd_after[k].numpy()) for k in keys])
    net.eval()
    with torch.no_grad():
        rec['out_after3'] = net(xt).numpy()
    return rec


def block_cases(arch_mod, common):
    """Per-block golden: forward, input-grad and param-grads for a fixed upstream gradient."""
    rs = np.random.RandomState(77)
    rec = {}

    def run(tag, mod, xshape):
        spec = [(k, tuple(v.shape)) for k, v in mod.state_dict().items()]
        sd = recipe.make_weights(spec, seed=sum(map(ord, tag)))
        mod.load_state_dict({k: t(v) for k, v in sd.items()})
        x = rs.uniform(-1, 1, size=xshape).astype(np.float32)
        xt = t(x).requires_grad_(True)
        out = mod(xt.clone())       # reference blocks use in-place `+=`; keep the leaf intact
        g = rs.uniform(-1, 1, size=tuple(out.shape)).astype(np.float32)
        out.backward(t(g))
        rec[tag + '::spec_keys'] = np.array([k for k, _ in spec])
        rec[tag + '::x'] = x
        rec[tag + '::g'] = g
        rec[tag + '::out'] = out.detach().numpy()
        rec[tag + '::dx'] = xt.grad.numpy()
        for k, p in mod.named_parameters():
            rec[tag + '::gradsub::' + k] = recipe.subsample(p.grad.numpy()).copy()
            rec[tag + '::gradsum::' + k] = checksum(p.grad.numpy())

    act = torch.nn.ReLU(True)
    run('calayer', arch_mod.CALayer(64, 16), (2, 64, 9, 11))
    run('rcab', arch_mod.RCAB(common.default_conv, 64, 3, 16, act=act), (2, 64, 9, 11))
    run('resgroup', arch_mod.ResidualGroup(common.default_conv, 64, 3, 16, act=act, res_scale=1, n_resblocks=2),
        (1, 64, 10, 7))
    run('resblock', common.ResBlock(common.default_conv, 64, 3, act=act, res_scale=0.1), (2, 64, 8, 13))
    run('upsampler2', common.Upsampler(common.default_conv, 2, 64, act=False), (1, 64, 6, 5))
    run('upsampler3', common.Upsampler(common.default_conv, 3, 64, act=False), (1, 64, 5, 4))
    run('upsampler4', common.Upsampler(common.default_conv, 4, 64, act=False), (1, 64, 4, 6))
    run('conv64', common.default_conv(64, 64, 3), (2, 64, 11, 18))
    return rec


def set5_case(arch_mod):
    """BASELINE.json configs[0]: EDSR-baseline x4 on Data/example_data/Set5 (LR random-blur PNGs)."""
    from PIL import Image
    lr_dir = os.path.join(REF, 'Data/example_data/Set5/lr_random_blur')
    hr_dir = os.path.join(REF, 'Data/example_data/Set5/hr')
    spec = recipe.edsr_spec(16, 64, 4)
    sd = recipe.make_weights(spec, seed=5)
    net = arch_mod.EDSR()
    net.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
    net.eval()
    rec = {}
    names = sorted(f for f in os.listdir(lr_dir) if f.endswith('.png'))
    rec['names'] = np.array(names)
    for f in names:
        lr = np.asarray(Image.open(os.path.join(lr_dir, f)).convert('RGB'))
        hr = np.asarray(Image.open(os.path.join(hr_dir, f)).convert('RGB'))
        x = (lr.astype(np.float32) / 255.0).transpose(2, 0, 1)[None]     # ToTensor() semantics
        with torch.no_grad():
            out = net(t(x)).numpy()
        rec[f + '::lr_u8'] = lr
        rec[f + '::hr_u8'] = hr
        rec[f + '::out_checksum'] = checksum(out)
        rec[f + '::out_crop'] = out[:, :, :32, :32].copy()
        rec[f + '::out_ds8'] = out[:, :, ::8, ::8].copy()
    return rec


def qrcan_cases():
    """Q-RCAN (meta-attention) forward through the reference's QRCAN module; 'modulate' attributes through the
    reference handler's own scale_qpi arithmetic (attention_manipulators/handlers.py:59-73)."""
    from rumpy.SISR.models.attention_manipulators import architectures as qarch
    rec = {}
    for name in recipe.QCASES:
        kw, has_q, sd, x, meta = recipe.qcase_tensors(name)
        net = qarch.QRCAN(**kw)
        assert list(net.state_dict().keys()) == list(sd.keys()), 'Q-RCAN key order mismatch vs reference'
        net.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
        net.eval()
        attrs = t(meta).unsqueeze(2).unsqueeze(3)
        if kw['style'] == 'modulate':
            base = np.linspace(0, 1, 64)
            scaled = attrs * (0.8 - (-0.2)) + (-0.2)
            rows = []
            for i in range(scaled.size(0)):
                mu = scaled[i].squeeze().numpy()
                rows.append(torch.from_numpy((1 / (np.sqrt(2 * np.pi) * 0.2)) *
                                             np.exp(-np.power(base - mu, 2.) / (2 * np.power(0.2, 2.)))).type(torch.float32))
            attrs = torch.stack(rows).unsqueeze(2).unsqueeze(3)
        with torch.no_grad():
            out = net(t(x), attrs)
        rec[name + '::attributes'] = attrs.numpy()
        rec[name + '::out'] = out.numpy()
        print(name, 'out', out.shape, float(out.abs().max()))
        if True:
            # training ('modulate': the attributes carry no gradient, the networks' own parameters do): L1 gradients of every parameter (q-layers included) and 3 Adam steps
            y = recipe.make_input((x.shape[0], 3, x.shape[2] * kw['scale'], x.shape[3] * kw['scale']),
                                  recipe.QCASES[name][3] + 1000)
            net.train()
            o = net(t(x), attrs)
            loss = torch.nn.L1Loss()(o, t(y))
            loss.backward()
            rec[name + '::loss'] = np.float32(loss.item())
            for k, p in net.named_parameters():
                rec[name + '::gradsub::' + k] = recipe.subsample(p.grad.numpy()).copy()
            opt = torch.optim.Adam(net.parameters(), lr=1e-4)
            losses = []
            for _ in range(3):
                o = net(t(x), attrs)
                l = torch.nn.L1Loss()(o, t(y))
                opt.zero_grad()
                l.backward()
                opt.step()
                losses.append(l.item())
            rec[name + '::train_losses'] = np.array(losses, dtype=np.float32)
            net.eval()
            with torch.no_grad():
                rec[name + '::out_after3'] = net(t(x), attrs).numpy()
    for name in recipe.QECASES:
        kw, has_q, sd, x, meta = recipe.qecase_tensors(name)
        net = qarch.QEDSR(**kw)
        assert list(net.state_dict().keys()) == list(sd.keys()), 'Q-EDSR key order mismatch vs reference'
        net.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
        net.eval()
        attrs = t(meta).unsqueeze(2).unsqueeze(3)
        with torch.no_grad():
            out = net(t(x), attrs)
        rec[name + '::attributes'] = attrs.numpy()
        rec[name + '::out'] = out.numpy()
        print(name, 'out', out.shape, float(out.abs().max()))
        # training: L1 gradients of every parameter (q-layers included) and 3 Adam steps
        y = recipe.make_input((x.shape[0], 3, x.shape[2] * kw['scale'], x.shape[3] * kw['scale']),
                              recipe.QECASES[name][3] + 1000)
        net.train()
        o = net(t(x), attrs)
        loss = torch.nn.L1Loss()(o, t(y))
        loss.backward()
        rec[name + '::loss'] = np.float32(loss.item())
        for k, p in net.named_parameters():
            rec[name + '::gradsub::' + k] = recipe.subsample(p.grad.numpy()).copy()
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)
        losses = []
        for _ in range(3):
            o = net(t(x), attrs)
            l = torch.nn.L1Loss()(o, t(y))
            opt.zero_grad()
            l.backward()
            opt.step()
            losses.append(l.item())
        rec[name + '::train_losses'] = np.array(losses, dtype=np.float32)
        net.eval()
        with torch.no_grad():
            rec[name + '::out_after3'] = net(t(x), attrs).numpy()
    return rec


def han_cases(arch_mod):
    rec = {}
    for name in recipe.HCASES:
        nb, scale, sd, x = recipe.hcase_tensors(name)
        net = arch_mod.HAN(n_resblocks=nb, scale=scale)
        assert list(net.state_dict().keys()) == list(sd.keys()), 'HAN key order mismatch vs reference'
        net.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
        net.eval()
        with torch.no_grad():
            rec[name + '::out'] = net(t(x)).numpy()
        print(name, rec[name + '::out'].shape, float(np.abs(rec[name + '::out']).max()))
        # training: L1 gradients of every parameter (attention modules included) and 3 Adam steps
        y = recipe.make_input((x.shape[0], 3, x.shape[2] * scale, x.shape[3] * scale), recipe.HCASES[name][4] + 1000)
        net.train()
        loss = torch.nn.L1Loss()(net(t(x)), t(y))
        loss.backward()
        rec[name + '::loss'] = np.float32(loss.item())
        for k, p in net.named_parameters():
            rec[name + '::gradsub::' + k] = recipe.subsample(p.grad.numpy()).copy()
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)
        losses = []
        for _ in range(3):
            l = torch.nn.L1Loss()(net(t(x)), t(y))
            opt.zero_grad()
            l.backward()
            opt.step()
            losses.append(l.item())
        rec[name + '::train_losses'] = np.array(losses, dtype=np.float32)
    return rec


def qhan_case():
    from rumpy.SISR.models.attention_manipulators import architectures as qarch
    kw, has_q, sd, x, meta = recipe.qhcase_tensors()
    net = qarch.QHAN(**kw)
    assert list(net.state_dict().keys()) == list(sd.keys()), 'Q-HAN key order mismatch vs reference'
    net.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
    attrs = t(meta).unsqueeze(2).unsqueeze(3)
    rec = {}
    net.eval()
    with torch.no_grad():
        rec['qhan::out'] = net(t(x), attrs).numpy()
    y = recipe.make_input((x.shape[0], 3, x.shape[2] * kw['scale'], x.shape[3] * kw['scale']), recipe.QHCASE['xseed'] + 1000)
    net.train()
    loss = torch.nn.L1Loss()(net(t(x), attrs), t(y))
    loss.backward()
    rec['qhan::loss'] = np.float32(loss.item())
    for k, p in net.named_parameters():
        rec['qhan::gradsub::' + k] = recipe.subsample(p.grad.numpy()).copy()
    print('qhan', rec['qhan::out'].shape, float(rec['qhan::loss']))
    return rec


def main():
    torch.set_num_threads(os.cpu_count())
    arch_mod, common = import_reference()
    for name in recipe.CASES:
        rec = net_case(arch_mod, name)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **rec)
        print(name, 'loss', rec['loss'], 'train', rec['train_losses'], 'out', rec['out'].shape)
    np.savez_compressed(os.path.join(HERE, 'blocks.npz'), **block_cases(arch_mod, common))
    np.savez_compressed(os.path.join(HERE, 'set5_edsr_baseline.npz'), **set5_case(arch_mod))
    np.savez_compressed(os.path.join(HERE, 'qrcan.npz'), **qrcan_cases())
    np.savez_compressed(os.path.join(HERE, 'han.npz'), **{**han_cases(arch_mod), **qhan_case()})
    print('done')


if __name__ == '__main__':
    main()
