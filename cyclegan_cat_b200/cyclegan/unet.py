"""U-Net builders -- same names, config keys and error behaviour as the reference
`cyclegan/unet.py`, emitting the layer-graph IR that libcyclegan_b200.so executes.

`double_conv`      mirrors unet.py:20-36
`strided_unet`     mirrors unet.py:39-78
`unet_generator`   mirrors unet.py:81-124
Missing mandatory keys raise KeyError at build time (unittests/test_unet.py:41-72).
"""
from typing import Dict

from .. import ir
from ..runtime import Model


def _norm(g: ir.Graph, x, norm_type: str):
    """strided_unet's choice (unet.py:55-58,69-72): 'instancenorm' or, for anything else, BatchNormalization()."""
    if norm_type == 'instancenorm':
        return g.instance_norm(x, affine=True)          # TFA default center=scale=True, eps 1e-3
    return g.batch_norm(x, affine=True)                 # keras defaults: momentum .99, eps 1e-3, center=scale=True


def double_conv(g: ir.Graph, x, filter: int, kernel_size: int,
                norm_type: str = 'instancenorm', apply_dropout: bool = False):
    """Two (Conv k s1 SAME no-bias -> norm -> ReLU [-> Dropout(0.5)]) stages, unet.py:20-36.  As in the reference a
    normalization string that is neither 'batchnorm' nor 'instancenorm' (case-insensitive) adds no norm layer."""
    for _ in range(2):
        x = g.conv(x, filter, kernel_size, stride=1, padding='same', use_bias=False)
        if norm_type.lower() == 'batchnorm':
            x = g.batch_norm(x, affine=True)
        elif norm_type.lower() == 'instancenorm':
            x = g.instance_norm(x, affine=True)
        x = g.act(x, ir.ACT_RELU)
        if apply_dropout:
            x = g.dropout(x, 0.5)
    return x


def strided_unet(config: Dict, mode: str = "bf16") -> Model:
    filters = config['filters']
    kernel_sizes = config['kernels']
    norm_type = config['normalization']
    output_channels = config['output_channels']
    final_activation = config['final_activation']

    g = ir.Graph()
    skips = []
    x = g.input

    down_filters = filters
    up_filters = filters[::-1][:-1]
    for filter, kernel_size in list(zip(down_filters, kernel_sizes))[:-1]:
        x = g.conv(x, filter, kernel_size, stride=2, padding='same')
        x = _norm(g, x, norm_type)
        x = g.act(x, ir.ACT_RELU)
        skips.insert(0, x)

    x = g.conv(x, filters[-1], kernel_sizes[-1], stride=2, padding='same')

    for filter, skip, kernel_size in zip(up_filters, skips, kernel_sizes[:0:-1]):
        x = g.conv_transpose(x, filter, kernel_size, stride=2)
        x = g.concat(skip, x)
        x = _norm(g, x, norm_type)
        x = g.act(x, ir.ACT_RELU)

    x = g.conv_transpose(x, output_channels, 4, stride=2)
    x = g.act(x, ir.ACT_BY_NAME[final_activation])
    return Model(g, name="strided_unet", mode=mode)


def unet_generator(config: Dict, mode: str = "bf16") -> Model:
    filters = config['filters']
    kernel_sizes = config['kernels']
    expansion = config['expansion']
    norm_type = config['normalization']
    apply_dropout = config['dropout']
    output_channels = config['output_channels']
    final_activation = config['final_activation']

    g = ir.Graph()
    skips = []
    x = g.input

    down_filters = filters
    up_filters = filters[::-1][:-1]
    for filter, kernel_size in list(zip(down_filters, kernel_sizes))[:-1]:
        x = double_conv(g, x, filter, kernel_size, norm_type, apply_dropout)
        skips.insert(0, x)
        x = g.avg_pool(x)

    x = double_conv(g, x, down_filters[-1], kernel_sizes[-1], norm_type, apply_dropout)

    for filter, skip, kernel_size in zip(up_filters, skips, kernel_sizes[:0:-1]):
        if expansion == 'upsample':
            x = g.upsample(x)
        else:
            # the reference branch (unet.py:110-117) ends in `ReLU(x)`, which constructs a
            # layer from a tensor and fails at build time; keep that behaviour visible.
            raise TypeError("expansion != 'upsample' cannot build in the reference (unet.py:117)")
        x = g.concat(skip, x)
        x = double_conv(g, x, filter, kernel_size, norm_type, apply_dropout)

    x = g.conv(x, output_channels, 1, stride=1, padding='same', init="glorot")
    x = g.act(x, ir.ACT_BY_NAME[final_activation])
    return Model(g, name="unet_generator", mode=mode)
