"""Generates tests/golden/bicubic.npz by running the UNMODIFIED reference's bicubic baseline
(`EvalHub._low_res_prep`, rumpy/shared_framework/evaluation/standard_eval.py:240-275) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden_bicubic.py

The method is called unbound with a two-attribute stand-in for `self` (`data_type`, `scale` are all it reads), so
the arithmetic is the reference's own: torchvision ToPILImage -> PIL resize(BICUBIC) -> ToTensor.  Inputs are the
Set5 LR crops already stored in set5_edsr_baseline.npz plus seeded random / extreme-valued images; only the
reference's OUTPUTS (as uint8: they are k/255 exactly) are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402  (import shims)


def cases():
    rs = np.random.RandomState(11)
    out = {}
    for name, (n, c, h, w), scale in [('rand_x4', (2, 3, 19, 27), 4), ('rand_x2', (1, 3, 33, 8), 2),
                                      ('rand_x3', (1, 1, 7, 50), 3), ('tiny_x4', (1, 3, 1, 2), 4),
                                      ('rand_x8', (1, 3, 9, 9), 8)]:
        out[name] = (rs.rand(n, c, h, w).astype(np.float32), scale)
    out['extremes_x4'] = ((rs.rand(1, 3, 24, 40) > 0.5).astype(np.float32), 4)      # over / undershoot -> byte clamp
    gold = np.load(os.path.join(HERE, 'set5_edsr_baseline.npz'))
    f = str(gold['names'][0])
    out['set5_x4'] = ((gold[f + '::lr_u8'].astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)[None], 4)
    return out


def main():
    make_golden.import_reference()
    from rumpy.shared_framework.evaluation.standard_eval import EvalHub
    store = {}
    names = []
    for name, (x, scale) in cases().items():
        me = types.SimpleNamespace(data_type='single-frame', scale=scale)
        up, _ = EvalHub._low_res_prep(me, torch.from_numpy(x), timing=False)
        up = up.numpy()
        u8 = np.rint(up * 255.0).astype(np.uint8)
        assert np.array_equal(u8.astype(np.float32) / np.float32(255.0), up)       # stored losslessly as bytes
        store[name + '::lr'] = x
        store[name + '::scale'] = np.int32(scale)
        store[name + '::up_u8'] = u8
        names.append(name)
    store['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'bicubic.npz'), **store)
    print('wrote bicubic.npz:', names)


if __name__ == '__main__':
    main()
