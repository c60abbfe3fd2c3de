#!/usr/bin/env python
"""Small driver for ncu: runs the residual-trunk convolution (3x3, 256->256, 64x64, N images) forward + backward a few
times through the C-ABI, so that `ncu -k regex:conv_tc_kernel|wgrad_tc_kernel` captures exactly the dominant kernels
at the C3 shape.    python tools/prof_conv.py [N=32] [iters=3]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cyclegan_cat_b200 import ir  # noqa: E402
from cyclegan_cat_b200.runtime import Model  # noqa: E402
from tests.test_gpu_parity import _net_grads  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    g = ir.Graph(channels=[256])
    x = g.reflect_pad(g.input, 1)
    x = g.conv(x, 256, 3, stride=1, padding='valid')
    x = g.instance_norm(x, affine=False)
    x = g.act(x, ir.ACT_RELU)
    m = Model(g, name="trunk_conv", mode="bf16", seed=0)
    rng = np.random.RandomState(0)
    xin = rng.uniform(-1, 1, (n, 64, 64, 256)).astype(np.float32)
    dy = rng.normal(0, 1, (n, 64, 64, 256)).astype(np.float32)
    for _ in range(iters):
        y, dx, grads = _net_grads(m, xin, dy)
    print("ok", y.shape, float(np.abs(y).mean()))


if __name__ == "__main__":
    main()
