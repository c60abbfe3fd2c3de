"""GPU tests of the tcgen05 convolution kernels (forward, data gradient, weight gradient) against
(a) the CUDA-core kernels of the same library on bf16-representable weights (same arithmetic up to summation
order) and (b) the fp64 oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch

from cyclegan_cat_b200 import _lib, ir
from cyclegan_cat_b200.cyclegan.model import create_model
from cyclegan_cat_b200.runtime import Model, _ptr, _stream_ptr
from oracle import models as om, tf_ops as T
from tests import common as C
from tests.test_gpu_parity import _check_grads, _net_grads

pytestmark = pytest.mark.gpu


def _bf16_round(a):
    return torch.from_numpy(np.asarray(a, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def _block_net(cin, cout):
    """RPad1 -> Conv3x3 valid -> IN -> ReLU -> RPad1 -> Conv3x3 valid -> IN : the residual branch of resnet.py:26-34."""
    g = ir.Graph(channels=[cin])
    x = g.reflect_pad(g.input, 1)
    x = g.conv(x, cout, 3, stride=1, padding='valid')
    x = g.instance_norm(x, affine=False)
    x = g.act(x, ir.ACT_RELU)
    x = g.reflect_pad(x, 1)
    x = g.conv(x, cin, 3, stride=1, padding='valid')
    x = g.instance_norm(x, affine=False)
    return g


def _updown_net(cin, cout, k):
    """Conv k s2 same -> IN -> ReLU -> ConvT k s2 same -> IN: the down/up-sampling layers of resnet.py:49-60 and the
    strided U-Net / discriminator stride-2 stacks (unet.py:54,66; resnet.py:96)."""
    g = ir.Graph(channels=[cin])
    x = g.conv(g.input, cout, k, stride=2, padding='same')
    x = g.instance_norm(x, affine=True)
    x = g.act(x, ir.ACT_RELU)
    x = g.conv_transpose(x, cin, k, stride=2)
    x = g.instance_norm(x, affine=False)
    return g


def _stem_head_net(f):
    """RPad3 -> Conv7x7(3->f) -> IN -> ReLU -> RPad3 -> Conv7x7(f->3) -> tanh: the two image-side layers of
    resnet_generator (resnet.py:38-46, 68, 82)."""
    g = ir.Graph(channels=[3])
    x = g.reflect_pad(g.input, 3)
    x = g.conv(x, f, 7, stride=1, padding='valid')
    x = g.instance_norm(x, affine=False)
    x = g.act(x, ir.ACT_RELU)
    x = g.reflect_pad(x, 3)
    x = g.conv(x, 3, 7, stride=1, padding='valid')
    x = g.act(x, ir.ACT_TANH)
    return g


def _double_conv_net(cin, cout, k):
    """unet.py:20-36 double_conv: 2 x (Conv k s1 'same' no bias -> affine IN -> ReLU); channel counts are multiples of 16."""
    g = ir.Graph(channels=[cin])
    x = g.input
    for _ in range(2):
        x = g.conv(x, cout, k, stride=1, padding='same', use_bias=False)
        x = g.instance_norm(x, affine=True)
        x = g.act(x, ir.ACT_RELU)
    return g


def _make(graph, tc, seed=0):
    os.environ["CG_DISABLE_TC"] = "0" if tc else "1"
    try:
        m = Model(graph, name="block", mode="bf16", seed=seed)
        m.handle()
    finally:
        os.environ.pop("CG_DISABLE_TC", None)
    return m


@pytest.mark.parametrize("cin,cout,h,w,n", [(128, 128, 16, 16, 2), (256, 256, 64, 64, 1), (128, 256, 16, 32, 3),
                                            (256, 128, 8, 128, 1)])
def test_tc_conv_matches_cuda_core_conv(cin, cout, h, w, n):
    g = _block_net(cin, cout)
    a, b = _make(g, True), _make(g, False)
    rng = np.random.RandomState(1)
    ws = [_bf16_round(v + (rng.normal(0, 0.05, v.shape) if v.ndim == 1 else 0)) for v in a.get_weights()]
    a.set_weights(ws)
    b.set_weights(ws)
    x = _bf16_round(rng.uniform(-1, 1, (n, h, w, cin)))
    dy = _bf16_round(rng.normal(0, 1, (n, h, w, cin)))
    ya, dxa, ga = _net_grads(a, x, dy)
    yb, dxb, gb = _net_grads(b, x, dy)
    assert C.rel_l2(ya, yb) <= 4e-3, C.rel_l2(ya, yb)
    assert C.rel_l2(dxa, dxb) <= 2e-2, C.rel_l2(dxa, dxb)
    scale = max(np.linalg.norm(v) for v in gb)
    for i, (u, v) in enumerate(zip(ga, gb)):
        e = np.linalg.norm(u - v) / max(np.linalg.norm(v), 0.02 * scale)
        assert e <= 2e-2, (i, u.shape, e)


@pytest.mark.parametrize("cin,cout,k,h,w,n", [(64, 128, 3, 32, 32, 2), (128, 256, 4, 32, 64, 1), (256, 512, 4, 16, 16, 2),
                                              (128, 64, 3, 256, 256, 1), (64, 128, 4, 64, 64, 3)])
def test_tc_strided_and_transposed_convs_match_cuda_core(cin, cout, k, h, w, n):
    g = _updown_net(cin, cout, k)
    a, b = _make(g, True), _make(g, False)
    rng = np.random.RandomState(3)
    ws = [_bf16_round(v + (rng.normal(0, 0.05, v.shape) if v.ndim == 1 else 0)) for v in a.get_weights()]
    a.set_weights(ws)
    b.set_weights(ws)
    x = _bf16_round(rng.uniform(-1, 1, (n, h, w, cin)))
    dy = _bf16_round(rng.normal(0, 1, (n, h, w, cin)))
    ya, dxa, ga = _net_grads(a, x, dy)
    yb, dxb, gb = _net_grads(b, x, dy)
    assert C.rel_l2(ya, yb) <= 4e-3, C.rel_l2(ya, yb)
    assert C.rel_l2(dxa, dxb) <= 2e-2, C.rel_l2(dxa, dxb)
    scale = max(np.linalg.norm(v) for v in gb)
    for i, (u, v) in enumerate(zip(ga, gb)):
        e = np.linalg.norm(u - v) / max(np.linalg.norm(v), 0.02 * scale)
        assert e <= 2e-2, (i, u.shape, e)


@pytest.mark.parametrize("cin,cout,k,h,w,n", [(16, 16, 4, 64, 64, 2), (80, 32, 4, 32, 128, 1), (160, 64, 4, 32, 32, 2),
                                              (192, 128, 4, 16, 16, 3), (16, 32, 5, 128, 128, 1), (32, 64, 3, 64, 64, 2),
                                              (48, 32, 7, 16, 64, 1)])
def test_tc_unet_double_conv_matches_cuda_core(cin, cout, k, h, w, n):
    """16-channel-group (SWIZZLE_32B) mode of conv_tc_kernel: forward and data gradient of the U-Net convs, including the
    asymmetric 'same' padding of even kernels (k4 s1 -> (1,2)) that comes from TMA out-of-bounds fill."""
    g = _double_conv_net(cin, cout, k)
    a, b = _make(g, True), _make(g, False)
    rng = np.random.RandomState(6)
    ws = [_bf16_round(v + (rng.normal(0, 0.05, v.shape) if v.ndim == 1 else 0)) for v in a.get_weights()]
    a.set_weights(ws)
    b.set_weights(ws)
    x = _bf16_round(rng.uniform(-1, 1, (n, h, w, cin)))
    dy = _bf16_round(rng.normal(0, 1, (n, h, w, cout)))
    ya, dxa, ga = _net_grads(a, x, dy)
    yb, dxb, gb = _net_grads(b, x, dy)
    assert C.rel_l2(ya, yb) <= 4e-3, C.rel_l2(ya, yb)
    assert C.rel_l2(dxa, dxb) <= 5e-2, C.rel_l2(dxa, dxb)
    scale = max(np.linalg.norm(v) for v in gb)
    for i, (u, v) in enumerate(zip(ga, gb)):
        e = np.linalg.norm(u - v) / max(np.linalg.norm(v), 0.02 * scale)
        # a fraction p of ReLU masks flips between the arms (|y| below the bf16 rounding of y): relative error ~ sqrt(p),
        # the same bound as dx above; test_tc_wgrad16_matches_cuda_core is the tight check of the weight-gradient kernel
        assert e <= 5e-2, (i, u.shape, e)


@pytest.mark.parametrize("cin,cout,k,h,w,n", [(16, 16, 4, 64, 64, 2), (32, 16, 4, 128, 128, 1), (80, 32, 4, 32, 128, 1),
                                              (192, 128, 4, 16, 16, 3), (16, 32, 5, 16, 64, 2), (48, 32, 7, 16, 64, 1),
                                              (128, 64, 3, 32, 32, 2), (16, 16, 4, 256, 256, 1)])
def test_tc_wgrad16_matches_cuda_core(cin, cout, k, h, w, n):
    """wgrad16_tc_kernel (tap-stacked M, 16-channel SWIZZLE_32B boxes) in isolation: one conv -> affine IN -> ReLU, so both arms
    see the same dY up to the bf16 rounding of the conv output."""
    g = ir.Graph(channels=[cin])
    x_ = g.conv(g.input, cout, k, stride=1, padding='same', use_bias=False)
    x_ = g.instance_norm(x_, affine=True)
    x_ = g.act(x_, ir.ACT_LEAKY, 0.2)
    a, b = _make(g, True), _make(g, False)
    rng = np.random.RandomState(8)
    ws = [_bf16_round(v + (rng.normal(0, 0.05, v.shape) if v.ndim == 1 else 0)) for v in a.get_weights()]
    a.set_weights(ws)
    b.set_weights(ws)
    x = _bf16_round(rng.uniform(-1, 1, (n, h, w, cin)))
    dy = _bf16_round(rng.normal(0, 1, (n, h, w, cout)))
    ya, dxa, ga = _net_grads(a, x, dy)
    yb, dxb, gb = _net_grads(b, x, dy)
    assert C.rel_l2(ya, yb) <= 4e-3, C.rel_l2(ya, yb)
    assert C.rel_l2(dxa, dxb) <= 1e-2, C.rel_l2(dxa, dxb)
    assert ga[0].shape == (k, k, cin, cout)
    assert C.rel_l2(ga[0], gb[0]) <= 1e-2, C.rel_l2(ga[0], gb[0])


@pytest.mark.parametrize("cin,cout,k,h,w,n", [(3, 64, 4, 64, 64, 2), (3, 64, 4, 256, 256, 1), (3, 128, 3, 32, 64, 3), (4, 64, 4, 128, 32, 2)])
def test_tc_im2col_input_layer_matches_cuda_core(cin, cout, k, h, w, n):
    """The discriminators' input layer (resnet.py:96, Conv k s2 'same' on the image): receptive fields unfolded into one dense
    64-channel chunk, 1x1 tensor-core GEMM forward, tap-stacked weight gradient, GEMM + col2im data gradient."""
    g = ir.Graph(channels=[cin])
    x_ = g.conv(g.input, cout, k, stride=2, padding='same')
    x_ = g.instance_norm(x_, affine=False)
    x_ = g.act(x_, ir.ACT_LEAKY, 0.2)
    x_ = g.conv(x_, 2 * cout, 4, stride=2, padding='same')
    x_ = g.instance_norm(x_, affine=False)
    a, b = _make(g, True), _make(g, False)
    rng = np.random.RandomState(9)
    ws = [_bf16_round(v + (rng.normal(0, 0.05, v.shape) if v.ndim == 1 else 0)) for v in a.get_weights()]
    a.set_weights(ws)
    b.set_weights(ws)
    x = _bf16_round(rng.uniform(-1, 1, (n, h, w, cin)))
    dy = _bf16_round(rng.normal(0, 1, (n, h // 4, w // 4, 2 * cout)))
    ya, dxa, ga = _net_grads(a, x, dy)
    yb, dxb, gb = _net_grads(b, x, dy)
    assert C.rel_l2(ya, yb) <= 4e-3, C.rel_l2(ya, yb)
    assert C.rel_l2(dxa, dxb) <= 3e-2, C.rel_l2(dxa, dxb)
    scale = max(np.linalg.norm(v) for v in gb)
    for i, (u, v) in enumerate(zip(ga, gb)):
        e = np.linalg.norm(u - v) / max(np.linalg.norm(v), 0.02 * scale)
        assert e <= 3e-2, (i, u.shape, e)


@pytest.mark.parametrize("f,h,w,n", [(64, 64, 64, 2), (64, 32, 128, 3), (128, 256, 256, 1)])
def test_tc_stem_and_head_match_cuda_core(f, h, w, n):
    g = _stem_head_net(f)
    a, b = _make(g, True), _make(g, False)
    rng = np.random.RandomState(4)
    ws = [_bf16_round(v + (rng.normal(0, 0.05, v.shape) if v.ndim == 1 else 0)) for v in a.get_weights()]
    a.set_weights(ws)
    b.set_weights(ws)
    x = _bf16_round(rng.uniform(-1, 1, (n, h, w, 3)))
    dy = _bf16_round(rng.normal(0, 1, (n, h, w, 3)))
    ya, dxa, ga = _net_grads(a, x, dy)
    yb, dxb, gb = _net_grads(b, x, dy)
    # the head sums 7 bf16-rounded partial rows (S) instead of one fp32 accumulator: 2-3 extra roundings of 2^-9
    assert C.rel_l2(ya, yb) <= 8e-3, C.rel_l2(ya, yb)
    assert C.rel_l2(dxa, dxb) <= 2e-2, C.rel_l2(dxa, dxb)
    scale = max(np.linalg.norm(v) for v in gb)
    for i, (u, v) in enumerate(zip(ga, gb)):
        e = np.linalg.norm(u - v) / max(np.linalg.norm(v), 0.02 * scale)
        assert e <= 2e-2, (i, u.shape, e)


def test_tc_single_conv_against_fp64():
    """One conv, no norm: y = conv(reflect_pad(x)) + b, exact fp64 reference on bf16-representable data:
    the only errors are fp32 accumulation order and the bf16 rounding of the stored output."""
    cin, cout, h, w, n = 128, 64, 16, 16, 2
    g = ir.Graph(channels=[cin])
    x_ = g.reflect_pad(g.input, 1)
    x_ = g.conv(x_, cout, 3, stride=1, padding='valid')
    x_ = g.instance_norm(x_, affine=False)       # required consumer for the tensor-core path
    m = _make(g, True)
    rng = np.random.RandomState(2)
    wts = [_bf16_round(rng.normal(0, 0.05, v.shape)) for v in m.get_weights()]
    m.set_weights(wts)
    x = _bf16_round(rng.uniform(-1, 1, (n, h, w, cin)))
    xt = torch.from_numpy(x).double()
    ref = T.instance_norm(T.conv2d(T.reflection_pad(xt, 1, 1), torch.from_numpy(wts[0]).double(),
                                   torch.from_numpy(wts[1]).double(), 1, "valid")).numpy()
    y = m(x).numpy()
    assert C.rel_l2(y, ref) <= 6e-3, C.rel_l2(y, ref)


@pytest.mark.parametrize("filters,size", [(32, 64), (64, 64)])
def test_resnet_with_tc_layers_against_oracle(filters, size):
    """resnet_generator whose residual trunk (4f = 128 / 256 channels) runs on the tcgen05 kernels."""
    cfg = dict(type="resnet_generator", filters=filters)
    m = create_model(cfg, mode="bf16")
    o = om.create_model(cfg, torch.float64)
    w = om.init_variables(o.var_specs, 11)
    m.set_weights(w)
    o.load(w)
    rng = np.random.RandomState(5)
    x = rng.uniform(-1, 1, (2, size, size, 3)).astype(np.float32)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    yo = o.forward(xt)
    dy = rng.normal(0, 1, tuple(yo.shape)).astype(np.float32)
    ref = torch.autograd.grad(yo, [xt] + o.variables, torch.from_numpy(dy).double())
    y, dx, grads = _net_grads(m, x, dy)
    assert C.rel_l2(y, yo.detach().numpy()) <= 2e-2, C.rel_l2(y, yo.detach().numpy())
    assert C.rel_l2(dx, ref[0].numpy()) <= 0.6
    _check_grads(grads, [r.numpy() for r in ref[1:]], "bf16", "resnet_tc")
