"""Mirror of the reference's HAN attention modules (/root/reference/rumpy/SISR/models/advanced/HAN_blocks.py:8-76):
parameter containers with the reference's names (`gamma`, `conv.weight [1,1,3,3,3]`, `conv.bias`).  The arithmetic runs
in librumpy_b200.so (csrc/han.cu: lam_energy / lam_attention / lam_apply, csam_cat) as part of HAN.forward."""
import torch
from torch import nn

from rumpy_b200 import _lib


class LAM_Module(nn.Module):
    """Layer attention module (reference :8-41)."""

    def __init__(self, in_dim):
        super(LAM_Module, self).__init__()
        self.chanel_in = in_dim
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, x):
        raise _lib.RumpyB200Error('LAM_Module runs inside the native HAN forward only (no standalone / CPU path)')


class CSAM_Module(nn.Module):
    """Channel-spatial attention module (reference :44-76)."""

    def __init__(self, in_dim):
        super(CSAM_Module, self).__init__()
        self.chanel_in = in_dim
        self.conv = nn.Conv3d(1, 1, 3, 1, 1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        raise _lib.RumpyB200Error('CSAM_Module runs inside the native HAN forward only (no standalone / CPU path)')
