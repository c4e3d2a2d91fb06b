"""GPU tests of the eval glue and the device-side patch pipeline (csrc/glue.cu).
Integer / byte work: bit-exact against the numpy expressions the reference uses; PSNR within 1e-3 dB."""
import numpy as np
import pytest
import torch

import recipe

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('shape', [(1, 3, 1, 1), (2, 3, 17, 31), (3, 1, 64, 64), (1, 3, 231, 517), (2, 4, 9, 5)])
def test_quantize_u8_bit_exact_vs_numpy(shape):
    """`np.clip(x * 255, 0, 255).astype(np.uint8)` (truncation; reference sr_tools/visualization.py:31-61), including
    out-of-range values, exact k/255 grid points and values one ulp either side of them."""
    from rumpy_b200.shared_framework.data import quantize_u8_device
    rs = np.random.RandomState(5)
    x = rs.uniform(-0.2, 1.2, size=shape).astype(np.float32)
    flat = x.reshape(-1)
    grid = (np.arange(256, dtype=np.float32) / np.float32(255.0))
    k = min(flat.size // 3, 256)
    flat[:k] = grid[:k]
    flat[k:2 * k] = np.nextafter(grid[:k], np.float32(2.0))
    flat[2 * k:3 * k] = np.nextafter(grid[:k], np.float32(-1.0))
    want = np.clip(x * 255, 0, 255).astype(np.uint8).transpose(0, 2, 3, 1)
    got = quantize_u8_device(torch.from_numpy(x).to(DEV)).cpu().numpy()
    assert got.dtype == np.uint8 and got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize('shape', [(1, 3, 8, 8), (3, 3, 57, 86), (2, 3, 512, 384)])
def test_psnr_y_matches_host_definition(shape):
    from rumpy_b200.shared_framework.data import psnr_y, psnr_y_device
    rs = np.random.RandomState(6)
    hr = rs.uniform(0, 1, size=shape).astype(np.float32)
    sr = (hr + rs.normal(0, 0.05, size=shape)).astype(np.float32)        # leaves [0,1]: exercises the clip
    got = psnr_y_device(torch.from_numpy(sr).to(DEV), torch.from_numpy(hr).to(DEV)).cpu().numpy()
    for n in range(shape[0]):
        want = psnr_y(torch.from_numpy(sr[n:n + 1]), torch.from_numpy(hr[n:n + 1]))
        assert abs(float(got[n]) - want) <= 1e-3, (n, float(got[n]), want)
    same = psnr_y_device(torch.from_numpy(hr).to(DEV), torch.from_numpy(hr).to(DEV)).cpu().numpy()
    assert np.all(same == 100.0)                                          # metrics.py:41-42
    a = psnr_y_device(torch.from_numpy(sr).to(DEV), torch.from_numpy(hr).to(DEV))
    assert torch.equal(a, psnr_y_device(torch.from_numpy(sr).to(DEV), torch.from_numpy(hr).to(DEV)))   # deterministic


def test_psnr_y_set5_golden_within_tolerance(golden_dir):
    """PSNR(Y) of the reference's own Set5 outputs (golden crops) computed on the device equals the numpy oracle's."""
    import os
    from oracle import sr_numpy
    from rumpy_b200.shared_framework.data import psnr_y_device
    gold = np.load(os.path.join(golden_dir, 'set5_edsr_baseline.npz'))
    for f in [str(n) for n in gold['names']]:
        crop = gold[f + '::out_crop']                                     # 1 x 3 x 32 x 32 of the reference output
        hr = (gold[f + '::hr_u8'][:32, :32].astype(np.float32) / 255.0).transpose(2, 0, 1)[None]
        want = sr_numpy.psnr(sr_numpy.rgb_to_y(np.clip(crop, 0, 1)), sr_numpy.rgb_to_y(hr))
        got = float(psnr_y_device(torch.from_numpy(crop).to(DEV), torch.from_numpy(hr).to(DEV))[0])
        assert abs(got - want) <= 1e-3, (f, got, want)


@pytest.mark.parametrize('crop,scale,augment', [(16, 4, True), (7, 3, True), (24, 2, False), (48, 4, True)])
def test_device_patch_pipeline_bit_exact_vs_host_pipeline(crop, scale, augment):
    """DevicePairSet (one kernel per batch, images resident in HBM) against PairSet (numpy crop / flips / transpose +
    ToTensor on the host) with the same seed: identical tags and bit-identical fp32 batches."""
    from rumpy_b200.shared_framework.data import DevicePairSet, PairSet
    cfg = {'synthetic': 11, 'crop': crop, 'random_augment': augment}
    host, dev = PairSet(cfg, scale, seed=8), DevicePairSet(cfg, scale, seed=8, device=0)
    # ragged, non-square sources: replace the synthetic squares by rectangles of different sizes
    rs = np.random.RandomState(3)
    items = []
    for i in range(11):
        h, w = crop + rs.randint(0, 40), crop + rs.randint(0, 40)
        items.append((f'img_{i}', rs.randint(0, 256, (h, w, 3), dtype=np.uint8),
                      rs.randint(0, 256, (h * scale, w * scale, 3), dtype=np.uint8)))
    host.items = items
    dev2 = DevicePairSet.__new__(DevicePairSet)
    dev2.__dict__.update(dev.__dict__)
    dev2.items = items
    dev2._lr = [torch.from_numpy(a).to(DEV) for _, a, _ in items]
    dev2._hr = [torch.from_numpy(a).to(DEV) for _, _, a in items]
    dev2._lr_tab = torch.tensor([t.data_ptr() for t in dev2._lr], dtype=torch.int64, device=DEV)
    dev2._hr_tab = torch.tensor([t.data_ptr() for t in dev2._hr], dtype=torch.int64, device=DEV)
    n_batches = 0
    for _ in range(3):                                                    # three epochs: the RNG streams stay in step
        for hb, db in zip(host.batches(4), dev2.batches(4)):
            assert hb['tag'] == db['tag']
            assert torch.equal(hb['lr'], db['lr'].cpu()), 'LR patch batch differs'
            assert torch.equal(hb['hr'], db['hr'].cpu()), 'HR patch batch differs'
            n_batches += 1
    assert n_batches == 6


@pytest.mark.parametrize('name', ['square_x4', 'wide_x2', 'tall_x3_noaug', 'exact_fit_x4'])
def test_device_patch_pipeline_vs_reference_golden(name, golden_dir):
    """`rumpy_patch_batch` against patches cut by the reference's OWN `random_flip_rotate` + `image_patch_selection`
    (image_functions.py:287-362, called in data_handler.py:570-596's order; tests/golden/host_glue.npz written by
    make_golden_host.py): bit-exact for every sample of three passes over the set."""
    import os
    import make_golden_host as mgh
    from rumpy_b200.shared_framework.data import DevicePairSet
    gold = np.load(os.path.join(golden_dir, 'host_glue.npz'))
    cfg, scale, seed = mgh.PATCH_CASES[name]
    dev = DevicePairSet(cfg, scale, seed=seed, device=0)
    n = len(dev)
    k = 0
    for _ in range(mgh.PASSES):
        got = list(dev.batches(n, shuffle=False))
        assert len(got) == 1
        assert np.array_equal(got[0]['lr'].cpu().numpy(), gold[f'patch::{name}::lr'][k:k + n]), 'LR patches'
        assert np.array_equal(got[0]['hr'].cpu().numpy(), gold[f'patch::{name}::hr'][k:k + n]), 'HR patches'
        k += n


def test_bicubic_baseline_vs_reference_golden(golden_dir):
    """`rumpy_bicubic_upsample` against the reference's own `EvalHub._low_res_prep` outputs
    (evaluation/standard_eval.py:240-275; tests/golden/bicubic.npz): bit-exact."""
    import os
    from rumpy_b200.shared_framework.data import bicubic_upsample_device
    gold = np.load(os.path.join(golden_dir, 'bicubic.npz'))
    for name in [str(n) for n in gold['names']]:
        got = bicubic_upsample_device(torch.from_numpy(gold[name + '::lr']).to(DEV), int(gold[name + '::scale']))
        want = gold[name + '::up_u8'].astype(np.float32) / np.float32(255.0)
        assert tuple(got.shape) == want.shape and np.array_equal(got.cpu().numpy(), want), name


@pytest.mark.parametrize('shape,scale', [((1, 3, 1, 1), 2), ((2, 3, 37, 20), 4), ((1, 1, 5, 131), 3),
                                         ((3, 3, 48, 48), 4), ((1, 3, 65, 33), 2), ((1, 2, 17, 19), 8),
                                         ((16, 3, 64, 64), 4), ((1, 3, 9, 70), 5)])
def test_bicubic_baseline_bit_exact_vs_oracle(shape, scale):
    """Ragged shapes (tile edges in both directions, images smaller than the filter support), every scale, random and
    saturated content (over- / undershoot clamps in both passes), k/255 grid points +- 1 ulp in the quantiser."""
    from oracle import pil_resample
    from rumpy_b200.shared_framework.data import bicubic_upsample_device
    rs = np.random.RandomState(7)
    x = rs.rand(*shape).astype(np.float32)
    x[0, 0] = (rs.rand(*shape[2:]) > 0.5).astype(np.float32)
    flat = x[-1, -1].reshape(-1)
    grid = np.arange(256, dtype=np.float32) / np.float32(255.0)
    k = min(flat.size // 2, 256)
    flat[:k] = np.minimum(np.nextafter(grid[:k], np.float32(2.0)), np.float32(1.0))
    flat[k:2 * k] = np.maximum(np.nextafter(grid[:k], np.float32(-1.0)), np.float32(0.0))
    want = pil_resample.low_res_prep(x, scale)
    dev = torch.from_numpy(x).to(DEV)
    got = bicubic_upsample_device(dev, scale)
    assert np.array_equal(got.cpu().numpy(), want)
    assert torch.equal(got, bicubic_upsample_device(dev, scale))          # deterministic


def test_bicubic_baseline_1080p_frame_vs_pillow():
    """BASELINE configs[4]'s frame (1080 x 1920 -> 4320 x 7680) against Pillow itself, plus the size-independent
    properties: a constant image stays constant, and the result equals its own tiles' (any crop far from the border
    depends only on the 5 x 5 LR neighbourhood)."""
    from PIL import Image
    from rumpy_b200.shared_framework.data import bicubic_upsample_device
    rs = np.random.RandomState(9)
    img = rs.randint(0, 256, size=(1080, 1920, 3)).astype(np.uint8)
    x = torch.from_numpy(img.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0))[None].to(DEV)
    got = bicubic_upsample_device(x, 4)
    want = np.asarray(Image.fromarray(img).resize((7680, 4320), resample=Image.BICUBIC)).transpose(2, 0, 1)
    got_u8 = torch.round(got[0] * 255).to(torch.uint8).cpu().numpy()
    assert np.array_equal(got_u8, want)
    assert torch.equal(got[0].cpu(), torch.from_numpy(want.copy()).float().div(255))      # ToTensor (true division)
    const = bicubic_upsample_device(torch.full((1, 3, 100, 100), 77 / 255, device=DEV), 4)
    assert torch.equal(const, torch.full_like(const, 77 / 255))
    crop = bicubic_upsample_device(x[:, :, 500:540, 900:960].contiguous(), 4)
    assert torch.equal(crop[:, :, 16:-16, 16:-16], got[:, :, 2016:2144, 3616:3824])


@pytest.mark.parametrize('shape,scale', [((1, 3, 1, 1), 2), ((2, 3, 37, 20), 4), ((1, 1, 5, 131), 3),
                                         ((3, 3, 48, 48), 4), ((1, 3, 65, 33), 2), ((1, 2, 17, 19), 8), ((1, 3, 9, 70), 5)])
def test_lanczos_baseline_bit_exact_vs_oracle_and_pillow(shape, scale):
    """`rumpy_lanczos_upsample` (`--lanczos_upsample`, standard_eval.py:252-253) against the oracle's restatement of
    Pillow's Lanczos-3 resampler (pinned to Pillow in tests/test_oracle_golden.py) AND against Pillow itself: ragged
    shapes, images smaller than the filter support, every scale, saturated content."""
    from PIL import Image
    from oracle import pil_resample
    from rumpy_b200.shared_framework.data import lanczos_upsample_device
    rs = np.random.RandomState(11)
    x = rs.rand(*shape).astype(np.float32)
    x[0, 0] = (rs.rand(*shape[2:]) > 0.5).astype(np.float32)
    want = pil_resample.low_res_prep(x, scale, 'lanczos')
    dev = torch.from_numpy(x).to(DEV)
    got = lanczos_upsample_device(dev, scale)
    assert np.array_equal(got.cpu().numpy(), want)
    assert torch.equal(got, lanczos_upsample_device(dev, scale))          # deterministic
    u8 = pil_resample.to_u8(x[0, 0])
    pil = np.asarray(Image.fromarray(u8).resize((shape[3] * scale, shape[2] * scale), resample=Image.LANCZOS))
    assert np.array_equal(torch.round(got[0, 0] * 255).to(torch.uint8).cpu().numpy(), pil)


def test_lanczos_baseline_1080p_frame_vs_pillow():
    from PIL import Image
    from rumpy_b200.shared_framework.data import lanczos_upsample_device
    rs = np.random.RandomState(9)
    img = rs.randint(0, 256, size=(1080, 1920, 3)).astype(np.uint8)
    x = torch.from_numpy(img.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0))[None].to(DEV)
    got = lanczos_upsample_device(x, 4)
    want = np.asarray(Image.fromarray(img).resize((7680, 4320), resample=Image.LANCZOS)).transpose(2, 0, 1)
    assert torch.equal(got[0].cpu(), torch.from_numpy(want.copy()).float().div(255))      # ToTensor (true division)


def test_glue_rejects_cpu_tensors():
    from rumpy_b200 import _lib
    from rumpy_b200.shared_framework.data import bicubic_upsample_device, psnr_y_device, quantize_u8_device
    with pytest.raises(_lib.RumpyB200Error):
        bicubic_upsample_device(torch.rand(1, 3, 4, 4), 4)
    with pytest.raises(_lib.RumpyB200Error):
        bicubic_upsample_device(torch.rand(1, 3, 4, 4, device=DEV), 1)    # scale 2..8 only
    with pytest.raises(_lib.RumpyB200Error):
        quantize_u8_device(torch.rand(1, 3, 4, 4))
    with pytest.raises(_lib.RumpyB200Error):
        psnr_y_device(torch.rand(1, 3, 4, 4), torch.rand(1, 3, 4, 4))
