"""TEST INFRASTRUCTURE ONLY (see oracle/sr_numpy.py): numpy restatement of the bicubic baseline the reference's
evaluation computes next to every model ("bicubic baseline").

Reference call sites: `rumpy/shared_framework/evaluation/standard_eval.py:240-275` (`_low_res_prep`: per image
`ToPILImage()` -> `image.resize((w*scale, h*scale), resample=Image.BICUBIC)` -> `ToTensor()`) and
`rumpy/image_tools/image_manipulation/image_functions.py:38-41` (`upsample`).

The arithmetic lives in two third-party dependencies that are not under /root/reference (`requirements.txt`: pillow
and torchvision, both unpinned; this image: Pillow 12.2.0):
  * torchvision `to_pil_image`: float tensor -> `pic.mul(255).byte()` (fp32 multiply, truncation);
  * Pillow `Image.resize` on 8-bit bands = `ImagingResample` (src/libImaging/Resample.c): two passes, horizontal then
    vertical, each a per-output-pixel FIR with double-precision coefficients (Keys bicubic, a = -0.5, support 2,
    taps clipped to the image and re-normalised) rounded to 22-bit fixed point; accumulators start at 2^21, results
    are shifted right by 22 and clamped to a byte; the intermediate image is uint8;
  * torchvision `to_tensor`: uint8 -> fp32, `.div(255)`.
Pinned: tests/test_oracle_golden.py compares this file with Pillow itself (which IS installed in the image) on
random ragged shapes and scales, and with tests/golden/bicubic.npz generated through the reference's own function.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: coefficients of the 8-bit path are 22-bit fixed point
SUPPORT = 2.0                        # bicubic filter support


def bicubic_filter(x):
    """Keys cubic convolution kernel with a = -0.5, evaluated in the order Resample.c writes it (no FMA)."""
    a = -0.5
    x = np.abs(np.asarray(x, dtype=np.float64))
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def lanczos_filter(x):
    """Resample.c `lanczos_filter` (a = 3): sinc(x) * sinc(x / 3) on [-3, 3), libm `sin` element by element (numpy's
    vectorised sin is not guaranteed to round like libm's)."""
    import math

    def sinc(v):
        if v == 0.0:
            return 1.0
        v = v * math.pi
        return math.sin(v) / v
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    return np.array([sinc(float(v)) * sinc(float(v) / 3) if -3.0 <= v < 3.0 else 0.0 for v in x], dtype=np.float64)


FILTERS = {'bicubic': (bicubic_filter, SUPPORT), 'lanczos': (lanczos_filter, 3.0)}


def precompute_coeffs(in_size, out_size, resample='bicubic'):
    """Resample.c `precompute_coeffs` + `normalize_coeffs_8bpc` for the whole-image box.
    Returns (xmin[out], count[out], kk[out][ksize] int32)."""
    filter_fn, filter_support = FILTERS[resample]
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = filter_support * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xmin = np.zeros(out_size, dtype=np.int64)
    count = np.zeros(out_size, dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)           # C cast: truncation towards zero
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        w = filter_fn((np.arange(n, dtype=np.float64) + lo - center + 0.5) * ss)
        ww = 0.0
        for v in w:                                        # running sum in tap order, like the C loop
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        fixed = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS))
        kk[xx, :n] = fixed.astype(np.int64).astype(np.int32)   # (int) cast: truncation
        xmin[xx], count[xx] = lo, n
    return xmin, count, kk


def _pass_last_axis(img, out_size, resample='bicubic'):
    """One resampling pass along the last axis of a uint8 array."""
    xmin, count, kk = precompute_coeffs(img.shape[-1], out_size, resample)
    ksize = kk.shape[1]
    idx = np.minimum(xmin[:, None] + np.arange(ksize)[None, :], img.shape[-1] - 1)    # taps past `count` have k = 0
    acc = (img[..., idx].astype(np.int64) * kk.astype(np.int64)).sum(-1) + (1 << (PRECISION_BITS - 1))
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_u8(img, out_h, out_w, resample='bicubic'):
    """`Image.resize((out_w, out_h), Image.BICUBIC | Image.LANCZOS)` for uint8 arrays [..., H, W] (each band on its
    own)."""
    img = np.asarray(img, dtype=np.uint8)
    if out_w != img.shape[-1]:
        img = _pass_last_axis(img, out_w, resample)                        # horizontal pass first
    if out_h != img.shape[-2]:
        img = np.swapaxes(_pass_last_axis(np.swapaxes(img, -1, -2), out_h, resample), -1, -2)
    return img


def to_u8(x):
    """torchvision `to_pil_image` on a float tensor in [0, 1]: `pic.mul(255).byte()`."""
    return (np.asarray(x, dtype=np.float32) * np.float32(255.0)).astype(np.uint8)


def low_res_prep(lr, scale, upsample_function='bicubic'):
    """`EvalHub._low_res_prep(lr, upsample_function=...)`: N x C x H x W fp32 in [0,1] -> N x C x sH x sW.  'lanczos'
    (standard_eval.py:252-253) is restated here ahead of a device kernel for it (the product provides bicubic only)."""
    lr = np.asarray(lr, dtype=np.float32)
    up = resize_u8(to_u8(lr), lr.shape[-2] * scale, lr.shape[-1] * scale, upsample_function)
    return up.astype(np.float32) / np.float32(255.0)
