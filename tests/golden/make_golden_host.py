"""Generates tests/golden/host_glue.npz by running the UNMODIFIED reference's host-side helpers on CPU:

  * the training-patch transforms `random_flip_rotate` + `image_patch_selection`
    (rumpy/image_tools/image_manipulation/image_functions.py:287-362), called in the order
    `SuperResImages.image_augment_crop` calls them (rumpy/sr_tools/data_handler.py:570-596), on the synthetic uint8
    image pairs of `rumpy_b200.shared_framework.data.PairSet` -- pins SURVEY 8(f) rank 4 (host PairSet and the device
    kernel rumpy_patch_batch) against the reference instead of against the repo's own host code;
  * the meta-attention handlers' metadata glue `QModel.generate_channels`
    (rumpy/SISR/models/attention_manipulators/__init__.py:87-108) and `QRCANHandler.scale_qpi` / `gaussian`
    (attention_manipulators/handlers.py:57-73).

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden_host.py
Nothing from the reference is copied: it is imported, executed, and only its OUTPUTS are stored.
"""
from __future__ import annotations

import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden  # noqa: E402  (import shims for the reference)

PATCH_CASES = {
    # name: (PairSet cfg, scale, seed)
    'square_x4': ({'synthetic': 6, 'crop': 16, 'random_augment': True}, 4, 8),
    'wide_x2': ({'synthetic': 5, 'synthetic_hw': (20, 37), 'crop': 12, 'random_augment': True}, 2, 11),
    'tall_x3_noaug': ({'synthetic': 4, 'synthetic_hw': (33, 18), 'crop': 9, 'random_augment': False}, 3, 5),
    'exact_fit_x4': ({'synthetic': 4, 'synthetic_hw': (16, 16), 'crop': 16, 'random_augment': True}, 4, 3),
}
PASSES = 3      # every image is sampled this many times (consecutive draws of one generator)


def patch_cases():
    from rumpy.image_tools.image_manipulation.image_functions import image_patch_selection, random_flip_rotate
    from rumpy_b200.shared_framework.data import PairSet, to_tensor
    rec = {}
    for name, (cfg, scale, seed) in PATCH_CASES.items():
        ps = PairSet(cfg, scale, seed=seed)
        random.seed(seed)                      # the reference draws from the module-level generator
        lrs, hrs = [], []
        for _ in range(PASSES):
            for _, lr_u8, hr_u8 in ps.items:
                lr, hr = to_tensor(lr_u8), to_tensor(hr_u8)          # transforms.ToTensor() of the uint8 image
                if cfg['random_augment']:
                    lr, hr = random_flip_rotate(lr, hr, hflip=True, vflip=True, rot=True)
                lp, hp, _ = image_patch_selection(lr, crop_size=cfg['crop'], image_hr=hr, scale=scale, patch_type='random')
                lrs.append(torch.stack(lp, 0).squeeze().contiguous().numpy())
                hrs.append(torch.stack(hp, 0).squeeze().contiguous().numpy())
        rec[f'patch::{name}::lr'] = np.stack(lrs)
        rec[f'patch::{name}::hr'] = np.stack(hrs)
        print(name, rec[f'patch::{name}::lr'].shape, rec[f'patch::{name}::hr'].shape)
    return rec


def metadata_cases():
    from rumpy.SISR.models.attention_manipulators import QModel
    from rumpy.SISR.models.attention_manipulators.handlers import QRCANHandler
    rec = {}
    g = torch.Generator().manual_seed(8)
    qpi = torch.rand((7, 1, 1, 1), generator=g)
    for clamp in (False, True):
        h = types.SimpleNamespace(min_mu=-0.2, max_mu=0.8, base_scaler=np.linspace(0, 1, 64), clamp=clamp,
                                  gaussian=QRCANHandler.gaussian)
        rec[f'scale_qpi::clamp{int(clamp)}'] = QRCANHandler.scale_qpi(h, qpi).numpy()
    rec['scale_qpi::qpi'] = qpi.numpy()
    x = torch.zeros(4, 3, 8, 8)
    cases = {'two_of_three': (['blur_kernel', 'noise'], [('blur_kernel',), ('qpi',), ('noise',)], 2, (4, 3)),
             'single_key': (['qpi'], [('qpi',)], 1, (4,)),
             'all_keys': (['all'], [('a',), ('b',), ('c',)], 3, (4, 3))}
    for name, (wanted, keys, num, shape) in cases.items():
        meta = torch.rand(shape, generator=g)
        m = types.SimpleNamespace(num_metadata=num, metadata=wanted, style='standard')
        rec[f'channels::{name}::metadata'] = meta.numpy()
        rec[f'channels::{name}::out'] = QModel.generate_channels(m, x, meta, keys).numpy()
    return rec


def main():
    make_golden.import_reference()
    rec = {**patch_cases(), **metadata_cases()}
    np.savez_compressed(os.path.join(HERE, 'host_glue.npz'), **rec)
    print('wrote host_glue.npz', len(rec), 'arrays')


if __name__ == '__main__':
    main()
