"""CPU-only checks: the C-ABI library loads and exports every symbol the public header declares, the ctypes
binding covers them, the module mirrors keep the reference's state_dict layout, host-side logic (registry,
sharding, gradient all-reduce over gloo with world_size 2) works, and nothing computes without a GPU."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import recipe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'rumpy_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(rumpy_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_header_symbol():
    import ctypes
    from rumpy_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip('librumpy_b200.so not built (run __graft_entry__.build())')
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/rumpy_b200.h but not exported'
    assert sorted(_lib.SIGNATURES) == syms, 'ctypes SIGNATURES and header disagree'
    assert lib.rumpy_version() == 100
    # ... and nothing else: no undeclared entry points, no process-global debug switches (every knob is a per-handle
    # option, rumpy_net_set_option)
    import shutil
    import subprocess
    if shutil.which('nm'):
        out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
        exported = sorted(line.split()[-1] for line in out.splitlines()
                          if ' T ' in line and line.split()[-1].startswith('rumpy_'))
        assert exported == syms, f'exported but not declared: {sorted(set(exported) - set(syms))}'


def test_no_gpu_fails_loudly():
    from rumpy_b200 import _lib
    if torch.cuda.is_available() or not os.path.exists(_lib.LIB_PATH):
        pytest.skip('only meaningful on a CPU-only box with the library built')
    lib = _lib.load()
    assert lib.rumpy_device_check() != 0
    assert b'CUDA' in lib.rumpy_last_error() or b'device' in lib.rumpy_last_error()
    from rumpy_b200.SISR.models.advanced.architectures import RCAN
    with pytest.raises(_lib.RumpyB200Error):
        RCAN(n_resgroups=1, n_resblocks=1)(torch.rand(1, 3, 8, 8))


def test_glue_host_side_argument_checks_without_a_gpu():
    """Host-only parts of the eval-glue ABI (no kernel runs): workspace sizes, argument validation, and the loud
    failure of the compute entry points when there is no CUDA device."""
    import ctypes
    from rumpy_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip('librumpy_b200.so not built (run __graft_entry__.build())')
    lib = _lib.load()
    # bicubic: OW column tap entries of 32 B + ceil(OH / 4) row-group entries of 144 B; scale 2..8 only
    assert lib.rumpy_bicubic_workspace(48, 48, 4) == 192 * 32 + 48 * 144
    assert lib.rumpy_bicubic_workspace(5, 7, 3) == 21 * 32 + 4 * 144
    assert lib.rumpy_bicubic_workspace(48, 48, 1) < 0 and lib.rumpy_bicubic_workspace(48, 48, 9) < 0
    assert lib.rumpy_bicubic_workspace(0, 48, 4) < 0
    assert lib.rumpy_psnr_y_workspace(3) > 0 and lib.rumpy_psnr_y_workspace(0) < 0
    assert lib.rumpy_bicubic_upsample(None, None, None, 1, 3, 8, 8, 4, None) != 0          # null pointers
    assert b'bicubic_upsample' in lib.rumpy_last_error()
    if not torch.cuda.is_available():
        buf = (ctypes.c_float * 16)()
        p = ctypes.cast(buf, ctypes.POINTER(ctypes.c_float))
        assert lib.rumpy_bicubic_upsample(p, p, ctypes.cast(buf, ctypes.c_void_p), 1, 1, 2, 2, 2, None) != 0
        from rumpy_b200.shared_framework.data import bicubic_upsample_device
        with pytest.raises(_lib.RumpyB200Error):
            bicubic_upsample_device(torch.rand(1, 3, 4, 4), 4)


def test_best_model_selection_criteria(tmp_path):
    """`BaseModel.best_model_selection_criteria` (reference base_architecture.py:601-612, helper_functions.py:29-40):
    first row with the highest val-PSNR / lowest loss, from a DataFrame, a dict of lists or the summary.csv itself."""
    import pandas as pd
    from rumpy_b200.shared_framework.models.base_architecture import BaseModel
    stats = {'epoch': [0, 1, 2, 3], 'train-loss': [0.5, 0.2, 0.3, 0.2], 'val-PSNR': [20.0, 27.5, 27.5, 25.0]}
    assert BaseModel.best_model_selection_criteria(stats=stats) == 1
    assert BaseModel.best_model_selection_criteria(stats=pd.DataFrame(stats), base_metric='train-loss') == 1
    assert BaseModel.best_model_selection_criteria(stats=pd.DataFrame(stats)) == int(pd.DataFrame(stats)['val-PSNR'].idxmax())
    pd.DataFrame(stats).to_csv(tmp_path / 'summary.csv', index=False)
    assert BaseModel.best_model_selection_criteria(log_dir=str(tmp_path)) == 1
    assert BaseModel.best_model_selection_criteria(stats_dir=str(tmp_path), base_metric='train-loss') == 1
    with pytest.raises(KeyError):
        BaseModel.best_model_selection_criteria(stats=stats, base_metric='nonsense')


def test_state_dict_layout_matches_reference_spec():
    from rumpy_b200.SISR.models.advanced.architectures import RCAN, EDSR
    m = RCAN()
    spec = recipe.rcan_spec()
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in spec]
    assert [k for k, _ in m.named_parameters()] == [k for k, _ in spec]      # Adam state is index-keyed
    assert sum(p.numel() for p in m.parameters()) == 15592355
    e = EDSR(net_features=256, num_blocks=32, res_scale=0.1)
    assert sum(p.numel() for p in e.parameters()) == 43089923
    assert [(k, tuple(v.shape)) for k, v in EDSR().state_dict().items()] == \
        [(k, tuple(s)) for k, s in recipe.edsr_spec()]
    for scale in (2, 3, 4, 8):
        assert [k for k in RCAN(n_resgroups=1, n_resblocks=1, scale=scale).state_dict()] == \
            [k for k, _ in recipe.rcan_spec(1, 1, scale=scale)]


def test_qrcan_state_dict_layout_matches_reference_spec():
    """Q-RCAN mirrors keep the reference's REGISTRATION order (final_body first; attention and q-node before the
    convolutions inside a block): tests/golden/make_golden.py asserted recipe.qrcan_spec against the reference."""
    from rumpy_b200.SISR.models.attention_manipulators.architectures import QRCAN
    for name in recipe.QCASES:
        kw, has_q, sd, x, meta = recipe.qcase_tensors(name)
        m = QRCAN(**kw)
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, v.shape) for k, v in sd.items()], name
        assert m._cfg['block_has_q'] == has_q
    full = QRCAN(style='standard', num_metadata=10, include_q_layer=True)     # sample q-rcan.toml configuration
    assert sum(p.numel() for p in full.parameters()) == 15592355 + 200 * (10 * 32 + 32 + 32 * 64 + 64)
    from rumpy_b200.SISR.models.attention_manipulators.architectures import QEDSR
    for name in recipe.QECASES:
        kw, has_q, sd, x, meta = recipe.qecase_tensors(name)
        m = QEDSR(**kw)
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, v.shape) for k, v in sd.items()], name
        assert m._cfg['block_has_q'] == has_q
    with pytest.raises(NotImplementedError):
        QRCAN(style='max_concat')
    with pytest.raises(NotImplementedError):
        QRCAN(style='standard', include_sft_layer=True)
    with pytest.raises(RuntimeError):
        QRCAN(style='standard', reduction=8)


def test_han_state_dict_layout_matches_reference_spec():
    from rumpy_b200.SISR.models.advanced.architectures import HAN
    for name in recipe.HCASES:
        nb, scale, sd, x = recipe.hcase_tensors(name)
        m = HAN(n_resblocks=nb, scale=scale)
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, v.shape) for k, v in sd.items()], name
    assert sum(p.numel() for p in HAN().parameters()) == 15592355 + 3 + 27 + (64 * 704 * 9 + 64) + (64 * 128 * 9 + 64)


def test_qhan_state_dict_layout_matches_reference_spec():
    from rumpy_b200.SISR.models.attention_manipulators.architectures import QHAN
    kw, has_q, sd, x, meta = recipe.qhcase_tensors()
    m = QHAN(**kw)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, v.shape) for k, v in sd.items()]
    assert m._cfg['block_has_q'] == has_q


def test_registry_and_legacy_switch():
    from rumpy_b200.shared_framework.models import available_models
    from rumpy_b200.shared_framework.models.base_architecture import BaseModel
    assert set(available_models) == {'rcan', 'edsr', 'han', 'qrcan', 'qedsr', 'qhan'}
    sd = {'model.module.head.0.weight': 1, 'model.body.0.bias': 2, 'tail.1.bias': 3}
    assert list(BaseModel.legacy_switch(sd)) == ['head.0.weight', 'body.0.bias', 'tail.1.bias']
    with pytest.raises(RuntimeError):
        from rumpy_b200.shared_framework.models import define_model
        define_model('rcan', device='cpu', model_save_dir='/tmp', eval_mode=True)


def test_patch_geometry_reproduces_host_pipeline():
    """`PairSet.geometry` (the 24-byte row the device patch kernel consumes) draws the same random numbers as
    `PairSet.sample`; applying its crop offsets and flag bits (hflip, vflip, transpose in that order) with numpy
    reproduces `sample` bit for bit -- the host-side half of the DevicePairSet contract."""
    from rumpy_b200.shared_framework.data import PairSet, to_tensor
    cfg = {'synthetic': 9, 'crop': 12, 'random_augment': True}
    a, b = PairSet(cfg, 3, seed=5), PairSet(cfg, 3, seed=5)
    seen = set()
    for rep in range(4):
        for i in range(len(a)):
            _, lr, hr = a.sample(i)
            idx, y, x, flags, lh, lw = b.geometry(i)
            seen.add(flags)
            _, L, H = b.items[idx]
            assert (lh, lw) == L.shape[:2]
            pl, ph = L[y:y + 12, x:x + 12], H[y * 3:(y + 12) * 3, x * 3:(x + 12) * 3]
            if flags & 1:
                pl, ph = pl[:, ::-1], ph[:, ::-1]
            if flags & 2:
                pl, ph = pl[::-1], ph[::-1]
            if flags & 4:
                pl, ph = pl.transpose(1, 0, 2), ph.transpose(1, 0, 2)
            assert torch.equal(lr, to_tensor(np.ascontiguousarray(pl))) and torch.equal(hr, to_tensor(np.ascontiguousarray(ph)))
    assert len(seen) >= 6            # the flag combinations actually occur


def test_shard_round_robin():
    from rumpy_b200.parallel import shard_round_robin
    items = list(range(10))
    parts = [shard_round_robin(items, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == items and parts[1] == [1, 5, 9]


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from rumpy_b200.parallel import GradAllReduce, shard_round_robin, broadcast_parameters
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:' + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
flat = torch.arange(10000, dtype=torch.float32) * (rank + 1)
ar = GradAllReduce(bucket_bytes=4096)
ar(flat)
assert ar.world_size == 2
assert torch.equal(flat, torch.arange(10000, dtype=torch.float32) * 3), 'bucketed all-reduce mismatch'
flat = torch.arange(10000, dtype=torch.float32) * (rank + 1)
ar.chunked(flat, [(None, 7000, 10000), (None, 2500, 7000), (None, 0, 2500)])     # ranges as the chunked backward hands them over
assert torch.equal(flat, torch.arange(10000, dtype=torch.float32) * 3), 'chunked all-reduce mismatch'
try:
    ar.chunked(flat, [(None, 7000, 10000), (None, 0, 2500)])
    raise SystemExit('a gap in the gradient ranges must be rejected')
except ValueError:
    pass
lin = torch.nn.Linear(4, 4)
broadcast_parameters(lin)
w = lin.weight.detach().clone(); dist.all_reduce(w); assert torch.allclose(w, lin.weight.detach() * 2)
assert shard_round_robin(list(range(7))) == list(range(7))[rank::2]
dist.destroy_process_group()
print('ok', rank)
'''


def test_gloo_world_size_2_allreduce(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_pool_rows_is_host_arithmetic():
    """rumpy_pool_rows needs no device: two partial rows per conv tile, 16 x 8 tiles at 64 channels (weights resident in
    shared memory), 8 x 16 above."""
    from rumpy_b200 import _lib
    lib = _lib.load()
    for H, W in ((48, 48), (37, 29), (1080, 1920), (1, 1)):
        assert lib.rumpy_pool_rows(H, W, 64) == 2 * ((H + 15) // 16) * ((W + 7) // 8)
        assert lib.rumpy_pool_rows(H, W, 128) == 2 * ((H + 7) // 8) * ((W + 15) // 16)
    assert lib.rumpy_pool_rows(0, 8, 64) == 0
