"""Minimal LR/HR patch data source for the CLI mirrors (the reference's full pipeline, rumpy/sr_tools/data_handler.py,
is out of scope -- SURVEY.md section 2 #12).  Same semantics for the keys the EDSR/RCAN path uses: `lr` / `hr` image
directories matched by file name, `crop` = LR patch side (HR patch = crop*scale, image_functions.py:320-326),
`random_augment` = hflip / vflip / transpose with p=0.5 each (:346-362).  `synthetic = N` yields N random pairs."""
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
            self.items = [(f'synthetic_{i}', g.randint(0, 256, (side, side, 3), dtype=np.uint8),
                           g.randint(0, 256, (side * scale, side * scale, 3), dtype=np.uint8))
                          for i in range(int(cfg['synthetic']))]
        else:
            names = sorted(f for f in os.listdir(cfg['lr']) if f.lower().endswith(('.png', '.jpg', '.bmp')))
            self.items = [(n, _load(os.path.join(cfg['lr'], n)), _load(os.path.join(cfg['hr'], n))) for n in names]
        if not self.items:
            raise RuntimeError('no LR/HR image pairs found')

    def __len__(self):
        return len(self.items)

    def sample(self, idx):
        name, lr, hr = self.items[idx]
        if self.crop:
            c, s = self.crop, self.scale
            y = self.rng.randint(0, lr.shape[0] - c)
            x = self.rng.randint(0, lr.shape[1] - c)
            lr = lr[y:y + c, x:x + c]
            hr = hr[y * s:(y + c) * s, x * s:(x + c) * s]
        if self.augment:
            if self.rng.random() < 0.5:
                lr, hr = lr[:, ::-1], hr[:, ::-1]
            if self.rng.random() < 0.5:
                lr, hr = lr[::-1], hr[::-1]
            if self.rng.random() < 0.5:
                lr, hr = lr.transpose(1, 0, 2), hr.transpose(1, 0, 2)
        return name, to_tensor(np.ascontiguousarray(lr)), to_tensor(np.ascontiguousarray(hr))

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
