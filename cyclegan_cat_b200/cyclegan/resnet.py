"""ResNet generator / simple CNN discriminator -- same names and config keys as
the reference `cyclegan/resnet.py`, emitting the IR libcyclegan_b200.so executes.

`ReflectionPadding2D`   mirrors resnet.py:11-23 (known answer: unittests/test_resnet.py:31-47)
`residual`              mirrors resnet.py:26-35
`conv7s1`               mirrors resnet.py:38-46
`downsample`/`upsample` mirror resnet.py:49-60
`resnet_generator`      mirrors resnet.py:63-85
`simple_discriminator`  mirrors resnet.py:87-105
"""
from typing import Dict

from .. import ir
from ..runtime import Model, reflection_pad_device


class ReflectionPadding2D:
    """Standalone layer object; calling it runs the CUDA reflect-pad kernel (any real/int dtype)."""

    def __init__(self, padding=(1, 1), **kwargs):
        self.padding = tuple(padding)

    def compute_output_shape(self, s):
        return (s[0], s[1] + 2 * self.padding[0], s[2] + 2 * self.padding[1], s[3])

    def __call__(self, x, mask=None):
        w_pad, h_pad = self.padding
        if w_pad != h_pad:
            raise NotImplementedError("only symmetric paddings are used by the reference (resnet.py:27,31,39)")
        return reflection_pad_device(x, h_pad)


def residual(g: ir.Graph, layer, filters):
    x = g.reflect_pad(layer, 1)
    x = g.conv(x, filters, 3, stride=1, padding='valid')
    x = g.instance_norm(x, affine=False)
    x = g.act(x, ir.ACT_RELU)

    x = g.reflect_pad(x, 1)
    x = g.conv(x, filters, 3, stride=1, padding='valid')
    x = g.instance_norm(x, affine=False)
    return g.add(layer, x)


def conv7s1(g: ir.Graph, layer_input, filters, final):
    x = g.reflect_pad(layer_input, 3)
    x = g.conv(x, filters, 7, stride=1, padding='valid')
    if final:
        x = g.act(x, ir.ACT_TANH)
    else:
        x = g.instance_norm(x, affine=False)
        x = g.act(x, ir.ACT_RELU)
    return x


def downsample(g: ir.Graph, layer, filters):
    x = g.conv(layer, filters, 3, stride=2, padding='same')
    x = g.instance_norm(x, affine=False)
    return g.act(x, ir.ACT_RELU)


def upsample(g: ir.Graph, layer, filters):
    x = g.conv_transpose(layer, filters, 3, stride=2)
    x = g.instance_norm(x, affine=False)
    return g.act(x, ir.ACT_RELU)


def resnet_generator(config: Dict, mode: str = "bf16") -> Model:
    filters = config['filters']
    g = ir.Graph()
    x = g.input
    x = conv7s1(g, x, filters, False)
    x = downsample(g, x, filters * 2)
    x = downsample(g, x, filters * 4)
    for _ in range(9):                      # nine hard-coded blocks, resnet.py:71-79
        x = residual(g, x, filters * 4)
    x = upsample(g, x, filters * 2)
    x = upsample(g, x, filters)
    x = conv7s1(g, x, 3, True)
    return Model(g, name="resnet_generator", mode=mode)


def simple_discriminator(config: Dict, mode: str = "bf16") -> Model:
    down_filters = config['filters']
    kernel_size = config['kernels']
    norm_type = config['normalization']
    g = ir.Graph()
    x = g.input
    for kernel, filter in zip(kernel_size, down_filters):
        x = g.conv(x, filter, kernel, stride=2, padding='same')
        if norm_type == 'instancenorm':
            x = g.instance_norm(x, affine=False)
        else:
            x = g.batch_norm(x, affine=False)           # resnet.py:100 BatchNormalization(center=False, scale=False)
        x = g.act(x, ir.ACT_LEAKY, slope=0.2)
    x = g.conv(x, 1, 1, stride=1, padding='same')
    return Model(g, name="simple_discriminator", mode=mode)
