"""Host-side glue against outputs of the UNMODIFIED reference (tests/golden/host_glue.npz, written by
tests/golden/make_golden_host.py in the build container): the training-patch transforms
(image_functions.py:287-362 in data_handler.py:570-596's order) and the meta-attention metadata helpers
(attention_manipulators/__init__.py:87-108, handlers.py:57-73).  Bit-exact: these are index permutations and a
float64 formula rounded once."""
import os
import types

import numpy as np
import pytest
import torch

import make_golden_host as mgh
from rumpy_b200.shared_framework.data import PairSet

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'host_glue.npz'))


@pytest.mark.parametrize('name', list(mgh.PATCH_CASES))
def test_pairset_sample_matches_reference_transforms(name):
    cfg, scale, seed = mgh.PATCH_CASES[name]
    ps = PairSet(cfg, scale, seed=seed)
    k = 0
    for _ in range(mgh.PASSES):
        for idx in range(len(ps)):
            _, lr, hr = ps.sample(idx)
            assert np.array_equal(lr.numpy(), GOLD[f'patch::{name}::lr'][k]), f'{name}: LR patch {k}'
            assert np.array_equal(hr.numpy(), GOLD[f'patch::{name}::hr'][k]), f'{name}: HR patch {k}'
            k += 1


def _patch_from_geometry(img_u8, y, x, flags, side):
    """numpy restatement of csrc/glue.cu patch_batch_kernel's addressing: cut first, then H, V, T on the patch"""
    p = img_u8[y:y + side, x:x + side]
    if flags & 1:
        p = p[:, ::-1]
    if flags & 2:
        p = p[::-1]
    if flags & 4:
        p = p.transpose(1, 0, 2)
    return (p.astype(np.float32) / 255.0).transpose(2, 0, 1)


@pytest.mark.parametrize('name', list(mgh.PATCH_CASES))
def test_geometry_rows_for_the_device_kernel_match_reference_transforms(name):
    """`PairSet.geometry` (what rumpy_patch_batch consumes: corner in the ORIGINAL image + flags applied to the
    patch) must describe the same patch the reference cuts from the AUGMENTED image."""
    cfg, scale, seed = mgh.PATCH_CASES[name]
    ps = PairSet(cfg, scale, seed=seed)
    c, k = cfg['crop'], 0
    for _ in range(mgh.PASSES):
        for idx in range(len(ps)):
            i, y, x, flags, lr_h, lr_w = ps.geometry(idx)
            _, lr_u8, hr_u8 = ps.items[i]
            assert (lr_h, lr_w) == lr_u8.shape[:2] and 0 <= y <= lr_h - c and 0 <= x <= lr_w - c
            assert np.array_equal(_patch_from_geometry(lr_u8, y, x, flags, c), GOLD[f'patch::{name}::lr'][k])
            assert np.array_equal(_patch_from_geometry(hr_u8, y * scale, x * scale, flags, c * scale),
                                  GOLD[f'patch::{name}::hr'][k])
            k += 1


@pytest.mark.parametrize('clamp', [False, True])
def test_scale_qpi_matches_reference(clamp):
    from rumpy_b200.SISR.models.attention_manipulators.handlers import QRCANHandler
    h = types.SimpleNamespace(min_mu=-0.2, max_mu=0.8, base_scaler=np.linspace(0, 1, 64), clamp=clamp,
                              gaussian=QRCANHandler.gaussian)
    got = QRCANHandler.scale_qpi(h, torch.from_numpy(GOLD['scale_qpi::qpi'])).numpy()
    assert got.dtype == np.float32 and np.array_equal(got, GOLD[f'scale_qpi::clamp{int(clamp)}'])


@pytest.mark.parametrize('name,wanted,keys,num', [
    ('two_of_three', ['blur_kernel', 'noise'], [('blur_kernel',), ('qpi',), ('noise',)], 2),
    ('single_key', ['qpi'], [('qpi',)], 1),
    ('all_keys', ['all'], [('a',), ('b',), ('c',)], 3)])
def test_generate_channels_matches_reference(name, wanted, keys, num):
    from rumpy_b200.SISR.models.attention_manipulators import QModel
    m = types.SimpleNamespace(num_metadata=num, metadata=wanted, style='standard')
    m._metadata_table = lambda *a: QModel._metadata_table(m, *a)
    meta = torch.from_numpy(GOLD[f'channels::{name}::metadata'])
    got = QModel.generate_channels(m, torch.zeros(4, 3, 8, 8), meta, keys)
    assert got.dtype == torch.float32 and np.array_equal(got.numpy(), GOLD[f'channels::{name}::out'])
    with pytest.raises(RuntimeError):
        QModel.generate_channels(m, torch.zeros(4, 3, 8, 8), None, keys)
