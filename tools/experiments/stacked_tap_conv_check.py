"""Design check (CPU, numpy) for DESIGN.md section 7, candidate 3: a 3x3 64->64 conv computed as
`D^T[c_out][pixel] = W . X^T` with TWO TAPS STACKED ALONG M (rows 0-63 = W_a, rows 64-127 = W_b, both multiplying the
same shifted activation window) so that a tcgen05 MMA with M = 128 can take the weights as its (TMEM) A operand.

Pixels are indexed linearly over a halo-padded tile (pitch P = width + 2).  For a pair (a, b) of tap offsets with
b = a + delta the lower accumulator half at column p holds tap a's term of output pixel p and the upper half holds tap
b's term of output pixel p - delta, so   out[q] = sum_groups  low_g[q] + up_g[q + delta_g].
Taps are grouped by delta: {(-P-1, -P+1), (-1, +1), (P-1, P+1)} with delta = 2 and {(-P, 0), (P, none)} with delta = P:
two accumulator groups, five stacked MMAs per K step instead of nine.  This script checks that bookkeeping against a
direct convolution and prints how many accumulator columns a tile needs.

    python tools/experiments/stacked_tap_conv_check.py
"""
import numpy as np


def direct_conv(x, w):
    """x [Cin][H][W], w [Cout][Cin][3][3] -> [Cout][H][W], zero padding, cross-correlation (nn.Conv2d)."""
    cin, h, wd = x.shape
    xp = np.zeros((cin, h + 2, wd + 2), dtype=np.float64)
    xp[:, 1:-1, 1:-1] = x
    out = np.zeros((w.shape[0], h, wd), dtype=np.float64)
    for ky in range(3):
        for kx in range(3):
            out += np.einsum('oc,chw->ohw', w[:, :, ky, kx], xp[:, ky:ky + h, kx:kx + wd])
    return out


def stacked_conv(x, w):
    cin, h, wd = x.shape
    cout = w.shape[0]
    P = wd + 2
    # linear halo-padded pixel axis with a margin of P + 1 zeros either side so every shifted window stays in range
    margin = P + 1
    npix = (h + 2) * P
    xl = np.zeros((cin, npix + 2 * margin), dtype=np.float64)
    xp = np.zeros((cin, h + 2, P), dtype=np.float64)
    xp[:, 1:-1, 1:-1] = x
    xl[:, margin:margin + npix] = xp.reshape(cin, -1)

    def tap(off):                                   # linear offset -> (ky, kx) weight slice
        ky, kx = divmod(off + P + 1, P)
        return w[:, :, ky, kx]

    groups = [(2, [(-P - 1, -P + 1), (-1, 1), (P - 1, P + 1)]), (P, [(-P, 0), (P, None)])]
    # output pixels of interest: interior of the padded tile; accumulator columns needed: [first, last + delta]
    first, last = P + 1, (h + 1) * P - 2
    out_lin = np.zeros((cout, npix), dtype=np.float64)
    cols_needed = []
    for delta, pairs in groups:
        n_cols = last + delta - first + 1
        cols_needed.append(n_cols)
        acc = np.zeros((2 * cout, n_cols), dtype=np.float64)          # one M = 128 accumulator (two halves)
        for a, b in pairs:
            a_op = np.concatenate([tap(a), tap(b) if b is not None else np.zeros_like(tap(a))], axis=0)   # [128][Cin]
            window = xl[:, margin + first + a: margin + first + a + n_cols]                               # B operand
            acc += a_op @ window                                                                           # the MMAs
        q = np.arange(first, last + 1)
        out_lin[:, q] += acc[:cout, q - first] + acc[cout:, q - first + delta]                            # epilogue
    return out_lin.reshape(cout, h + 2, P)[:, 1:-1, 1:-1], cols_needed


def main():
    rs = np.random.RandomState(0)
    for (h, wd) in [(8, 16), (5, 7), (16, 16)]:
        x = rs.randn(64, h, wd)
        w = rs.randn(64, 64, 3, 3)
        got, cols = stacked_conv(x, w)
        want = direct_conv(x, w)
        err = np.abs(got - want).max() / np.abs(want).max()
        useful = h * wd
        print(f'tile {h}x{wd}: rel. max error {err:.2e}; accumulator columns per group {cols} for {useful} output pixels '
              f'(column efficiency {useful / max(cols):.2f}); 5 stacked MMAs per K step instead of 9 (tap efficiency 0.90)')
        assert err < 1e-12


if __name__ == '__main__':
    main()
