"""Mirror of the reference's meta-attention block
(/root/reference/rumpy/SISR/models/attention_manipulators/q_layer.py:5-45).

`ParaCALayer` keeps the reference's constructor, layer sizing rule and parameter names
(`attribute_integrator.<i>.weight/bias`).  Inside a native QRCAN the block is never called as a torch module: the
engine evaluates sigmoid(FC2 relu(FC1 metadata)) for every RCAB in one kernel (csrc/misc_kernels.cuh
`q_scale_kernel`) and the trunk kernels multiply it into the channel-attention vector.
"""
from torch import nn

from rumpy_b200 import _lib


class ParaCALayer(nn.Module):
    def __init__(self, network_channels, num_metadata, nonlinearity=False, num_layers=2, dropout=False,
                 dropout_probability=None):
        super(ParaCALayer, self).__init__()
        layers = []
        multiplier = num_layers
        inputs = [num_metadata]
        for i in range(num_layers):
            if num_metadata > 15:       # reference q_layer.py:28-31: layer widths between the metadata size and C
                inputs.append((network_channels - num_metadata) // multiplier + num_metadata)
            else:
                inputs.append(network_channels // multiplier)
            layers.append(nn.Conv2d(inputs[i], inputs[i + 1], 1, padding=0, bias=True))
            if nonlinearity and multiplier != 1:
                layers.append(nn.ReLU(inplace=True))
            if dropout and multiplier != 1:
                layers.append(nn.Dropout(p=dropout_probability))
            multiplier -= 1
        layers.append(nn.Sigmoid())
        self.attribute_integrator = nn.Sequential(*layers)
        self.layer_sizes = inputs
        self.nonlinearity = nonlinearity

    def forward(self, x, attributes):
        raise _lib.RumpyB200Error('ParaCALayer runs inside the native QRCAN trunk only (no standalone / CPU path)')
