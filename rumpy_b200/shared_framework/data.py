"""Minimal LR/HR patch data source for the CLI mirrors (the reference's full pipeline, rumpy/sr_tools/data_handler.py,
is out of scope -- SURVEY.md section 2 #12).  Same semantics for the keys the EDSR/RCAN path uses: `lr` / `hr` image
directories matched by file name, `crop` = LR patch side (HR patch = crop*scale, image_functions.py:320-326),
`random_augment` = hflip / vflip / transpose with p=0.5 each (:346-362), drawn and applied BEFORE the patch is cut, like
the reference's `image_augment_crop` (sr_tools/data_handler.py:570-596; pinned by tests/golden/host_glue.npz).  `synthetic = N` yields N random pairs."""
from __future__ import annotations

import os
import random

import numpy as np
import torch


def _load(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert('RGB'), dtype=np.uint8)


def to_tensor(img_u8):
    return torch.from_numpy(img_u8.astype(np.float32) / 255.0).permute(2, 0, 1).contiguous()   # ToTensor()


class PairSet:
    def __init__(self, cfg, scale, seed=8):
        self.scale, self.crop = scale, cfg.get('crop')
        self.augment = bool(cfg.get('random_augment'))
        self.rng = random.Random(seed)
        if cfg.get('synthetic'):
            g = np.random.RandomState(seed)
            side = (self.crop or 48) * 2
            sh, sw = cfg.get('synthetic_hw', (side, side))       # LR image size (default: square, twice the crop)
            self.items = [(f'synthetic_{i}', g.randint(0, 256, (sh, sw, 3), dtype=np.uint8),
                           g.randint(0, 256, (sh * scale, sw * scale, 3), dtype=np.uint8))
                          for i in range(int(cfg['synthetic']))]
        else:
            names = sorted(f for f in os.listdir(cfg['lr']) if f.lower().endswith(('.png', '.jpg', '.bmp')))
            self.items = [(n, _load(os.path.join(cfg['lr'], n)), _load(os.path.join(cfg['hr'], n))) for n in names]
        if not self.items:
            raise RuntimeError('no LR/HR image pairs found')

    def __len__(self):
        return len(self.items)

    def _draw(self, lr_h, lr_w):
        """The random draws of one sample, in the reference's order (sr_tools/data_handler.py:570-596): first the
        augmentation coins of `random_flip_rotate` (image_functions.py:346-350: hflip, vflip, transpose, p = 0.5
        each, one `random()` per coin), then the patch corner of `random_patch_selection` (:287-294: `randint` for
        the row, then the column) ON THE AUGMENTED IMAGE -- a transposed image swaps the two ranges."""
        flags = 0
        if self.augment:
            flags |= 1 if self.rng.random() < 0.5 else 0
            flags |= 2 if self.rng.random() < 0.5 else 0
            flags |= 4 if self.rng.random() < 0.5 else 0
        h, w = 0, 0
        if self.crop:
            aug_h, aug_w = (lr_w, lr_h) if flags & 4 else (lr_h, lr_w)
            h = self.rng.randint(0, max(0, aug_h - self.crop))
            w = self.rng.randint(0, max(0, aug_w - self.crop))
        return flags, h, w

    def sample(self, idx):
        """One (name, LR, HR) training pair, built the way the reference builds it: augment the whole image pair,
        then cut the patch (LR at (h, w), HR at (h, w) * scale)."""
        name, lr, hr = self.items[idx]
        flags, h, w = self._draw(lr.shape[0], lr.shape[1])
        if flags & 1:
            lr, hr = lr[:, ::-1], hr[:, ::-1]
        if flags & 2:
            lr, hr = lr[::-1], hr[::-1]
        if flags & 4:
            lr, hr = lr.transpose(1, 0, 2), hr.transpose(1, 0, 2)
        if self.crop:
            c, s = self.crop, self.scale
            lr = lr[h:h + c, w:w + c]
            hr = hr[h * s:(h + c) * s, w * s:(w + c) * s]
        return name, to_tensor(np.ascontiguousarray(lr)), to_tensor(np.ascontiguousarray(hr))

    def geometry(self, idx):
        """The same draws as `sample` (same order, same count) without touching pixels, as the row
        `rumpy_patch_batch` consumes: [image index, y, x, flags (1 hflip | 2 vflip | 4 transpose, applied to the
        PATCH in that order), lr_h, lr_w] with (y, x) the patch corner in the ORIGINAL image.  Cutting at (h, w)
        after the flips equals cutting the mirrored window before them: undo the transpose (swap), then the
        vertical / horizontal mirror (corner -> size - crop - corner)."""
        _, lr, _ = self.items[idx]
        lr_h, lr_w = lr.shape[0], lr.shape[1]
        flags, h, w = self._draw(lr_h, lr_w)
        c = self.crop
        y, x = (w, h) if flags & 4 else (h, w)
        if flags & 2:
            y = lr_h - c - y
        if flags & 1:
            x = lr_w - c - x
        return [idx, y, x, flags, lr_h, lr_w]

    def batches(self, batch_size, shuffle=True, rank=0, world=1):
        order = list(range(len(self.items)))
        if shuffle:
            self.rng.shuffle(order)
        order = order[rank::world]
        for i in range(0, len(order) - batch_size + 1, batch_size):
            picks = [self.sample(j) for j in order[i:i + batch_size]]
            yield {'tag': [p[0] for p in picks], 'lr': torch.stack([p[1] for p in picks]),
                   'hr': torch.stack([p[2] for p in picks])}


def psnr_y(sr, hr, max_value=1.0):
    """PSNR on Y of jpg-style YCbCr after clipping to [0,1] (sr_tools/metrics.py:33-44,109-121;
    image_functions.py:72-88; base_interface.py:208-222).  sr, hr: NCHW float tensors on the CPU."""
    def y(img):
        img = img.clamp(0, 1)
        return 0.299 * img[:, 0] + 0.587 * img[:, 1] + 0.114 * img[:, 2]
    mse = torch.mean((y(sr.float()) - y(hr.float())) ** 2).item()
    if mse == 0:
        return 100.0
    return float(20 * np.log10(max_value / np.sqrt(mse)))


# ----------------------------------------------------------------------------------------------------------------
# Device-side versions (csrc/glue.cu): same results as the host functions above, computed where the batch already is
# ----------------------------------------------------------------------------------------------------------------
def psnr_y_device(sr, hr, max_value=1.0):
    """Per-image PSNR(Y) of NCHW fp32 CUDA tensors -> fp32 CUDA tensor [N] (no host sync).  Same definition as
    `psnr_y`, evaluated per image (what EvalHub records in individual_metrics.csv)."""
    from rumpy_b200 import _lib
    if not (sr.is_cuda and hr.is_cuda):
        raise _lib.RumpyB200Error('psnr_y_device: CUDA tensors only (no CPU fallback)')
    sr, hr = sr.contiguous().float(), hr.contiguous().float()
    if sr.shape != hr.shape or sr.dim() != 4 or sr.shape[1] != 3:
        raise ValueError(f'psnr_y_device: expected matching N x 3 x H x W tensors, got {tuple(sr.shape)} / {tuple(hr.shape)}')
    n, _, h, w = sr.shape
    lib = _lib.load()
    ws = torch.empty(lib.rumpy_psnr_y_workspace(n), dtype=torch.uint8, device=sr.device)
    out = torch.empty(n, dtype=torch.float32, device=sr.device)
    _lib.call('rumpy_psnr_y', sr.data_ptr(), hr.data_ptr(), out.data_ptr(), ws.data_ptr(), n, h, w, float(max_value),
              torch.cuda.current_stream().cuda_stream)
    return out


def quantize_u8_device(img):
    """NCHW fp32 CUDA tensor -> NHWC uint8 CUDA tensor, `np.clip(x * 255, 0, 255).astype(np.uint8)` (truncation)."""
    from rumpy_b200 import _lib
    if not img.is_cuda:
        raise _lib.RumpyB200Error('quantize_u8_device: CUDA tensors only (no CPU fallback)')
    img = img.contiguous().float()
    n, c, h, w = img.shape
    out = torch.empty((n, h, w, c), dtype=torch.uint8, device=img.device)
    _lib.call('rumpy_quantize_u8', img.data_ptr(), out.data_ptr(), n, c, h, w, torch.cuda.current_stream().cuda_stream)
    return out


def bicubic_upsample_device(lr, scale):
    """The evaluation's bicubic baseline on the device: `EvalHub._low_res_prep(lr, upsample_function='bicubic')`
    (reference evaluation/standard_eval.py:240-275) = per image ToPILImage -> `Image.resize(..., BICUBIC)` ->
    ToTensor.  N x C x H x W fp32 CUDA tensor in [0,1] -> N x C x sH x sW fp32 CUDA tensor, bit-exact with Pillow."""
    from rumpy_b200 import _lib
    if not lr.is_cuda:
        raise _lib.RumpyB200Error('bicubic_upsample_device: CUDA tensors only (no CPU fallback)')
    if lr.dim() != 4:
        raise ValueError(f'bicubic_upsample_device: expected an N x C x H x W tensor, got {tuple(lr.shape)}')
    lr = lr.contiguous().float()
    n, c, h, w = lr.shape
    ws_bytes = _lib.load().rumpy_bicubic_workspace(h, w, int(scale))
    if ws_bytes < 0:
        raise _lib.RumpyB200Error(f'bicubic_upsample_device: unsupported H={h} W={w} scale={scale} (scale 2..8)')
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=lr.device)
    out = torch.empty((n, c, h * int(scale), w * int(scale)), dtype=torch.float32, device=lr.device)
    _lib.call('rumpy_bicubic_upsample', lr.data_ptr(), out.data_ptr(), ws.data_ptr(), n, c, h, w, int(scale),
              torch.cuda.current_stream().cuda_stream)
    return out


def lanczos_upsample_device(lr, scale):
    """`EvalHub._low_res_prep(lr, upsample_function='lanczos')` (reference evaluation/standard_eval.py:252-253, the
    `--lanczos_upsample` switch) on the device: ToPILImage -> `Image.resize(..., LANCZOS)` -> ToTensor per image,
    bit-exact with Pillow.  N x C x H x W fp32 CUDA tensor in [0,1] -> N x C x sH x sW fp32 CUDA tensor."""
    from rumpy_b200 import _lib
    if not lr.is_cuda:
        raise _lib.RumpyB200Error('lanczos_upsample_device: CUDA tensors only (no CPU fallback)')
    if lr.dim() != 4:
        raise ValueError(f'lanczos_upsample_device: expected an N x C x H x W tensor, got {tuple(lr.shape)}')
    lr = lr.contiguous().float()
    n, c, h, w = lr.shape
    ws_bytes = _lib.load().rumpy_lanczos_workspace(h, w, int(scale))
    if ws_bytes < 0:
        raise _lib.RumpyB200Error(f'lanczos_upsample_device: unsupported H={h} W={w} scale={scale} (scale 2..8)')
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=lr.device)
    out = torch.empty((n, c, h * int(scale), w * int(scale)), dtype=torch.float32, device=lr.device)
    _lib.call('rumpy_lanczos_upsample', lr.data_ptr(), out.data_ptr(), ws.data_ptr(), n, c, h, w, int(scale),
              torch.cuda.current_stream().cuda_stream)
    return out


class DevicePairSet(PairSet):
    """PairSet whose uint8 images live in HBM and whose training batches (crop + flips + transpose + ToTensor) are
    produced by ONE kernel per batch (`rumpy_patch_batch`).  Draws the same random numbers in the same order as
    PairSet, so for one seed both yield bit-identical batches; only 24 B of geometry per sample cross PCIe."""

    def __init__(self, cfg, scale, seed=8, device=0):
        super().__init__(cfg, scale, seed)
        if not self.crop:
            raise ValueError('DevicePairSet needs `crop` (training patches); use PairSet for whole images')
        self.device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        self._lr = [torch.from_numpy(np.array(lr, copy=True)).to(self.device) for _, lr, _ in self.items]
        self._hr = [torch.from_numpy(np.array(hr, copy=True)).to(self.device) for _, _, hr in self.items]
        self._lr_tab = torch.tensor([t.data_ptr() for t in self._lr], dtype=torch.int64, device=self.device)
        self._hr_tab = torch.tensor([t.data_ptr() for t in self._hr], dtype=torch.int64, device=self.device)

    def batches(self, batch_size, shuffle=True, rank=0, world=1):
        from rumpy_b200 import _lib
        order = list(range(len(self.items)))
        if shuffle:
            self.rng.shuffle(order)
        order = order[rank::world]
        c, s = self.crop, self.scale
        for i in range(0, len(order) - batch_size + 1, batch_size):
            picks = order[i:i + batch_size]
            geom = torch.tensor([self.geometry(j) for j in picks], dtype=torch.int32).pin_memory()
            geom = geom.to(self.device, non_blocking=True)
            lr = torch.empty((batch_size, 3, c, c), dtype=torch.float32, device=self.device)
            hr = torch.empty((batch_size, 3, c * s, c * s), dtype=torch.float32, device=self.device)
            _lib.call('rumpy_patch_batch', self._lr_tab.data_ptr(), self._hr_tab.data_ptr(), geom.data_ptr(),
                      lr.data_ptr(), hr.data_ptr(), batch_size, c, s, torch.cuda.current_stream().cuda_stream)
            yield {'tag': [self.items[j][0] for j in picks], 'lr': lr, 'hr': hr}
