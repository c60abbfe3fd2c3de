"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on identical
weights and inputs.  Tolerances are BASELINE.json's: relative L2 <= 1e-4 in fp32 check mode,
<= 2e-2 in bf16 mode.  Gradients that are zero by construction (conv biases that feed an instance
norm, SURVEY.md 7 'Hard parts') are checked with an absolute bound scaled by the layer's kernel-
gradient norm instead."""
import os

import numpy as np
import pytest
import torch

from cyclegan_cat_b200.cyclegan.model import CycleGan, create_model
from cyclegan_cat_b200.cyclegan.resnet import ReflectionPadding2D, resnet_generator, simple_discriminator
from cyclegan_cat_b200.cyclegan.unet import strided_unet, unet_generator
from oracle import models as om
from oracle.train import OracleCycleGan, synthetic_batch
from tests import common as C

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "bf16": 2e-2}
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _pair(cfg, mode, seed=7):
    m = create_model(cfg, mode=mode)
    o = om.create_model(cfg)
    w = om.init_variables(o.var_specs, seed)
    # make IN gamma/beta and biases non-trivial so their paths are exercised
    rng = np.random.RandomState(seed + 1)
    w = [a + rng.normal(0, 0.05, a.shape).astype(np.float32) if a.ndim == 1 else a for a in w]
    m.set_weights(w)
    o.load(w)
    return m, o


# ---- the reference's own unit tests, run against the B200 builders ---------------------------------
def test_pooled_unet_shape():              # unittests/test_unet.py:27-31 (float64 numpy input accepted)
    assert unet_generator(C.FIX_UNET)(np.ones((1, 128, 128, 3))).shape == (1, 128, 128, 3)


def test_strided_unet_shape():             # unittests/test_unet.py:34-38
    assert strided_unet(C.FIX_UNET)(np.ones((1, 128, 128, 3))).shape == (1, 128, 128, 3)


def test_resnet_shape():                   # unittests/test_resnet.py:24-28
    assert resnet_generator(C.FIX_RESNET)(np.ones((1, 128, 128, 3))).shape == (1, 128, 128, 3)


def test_simple_discriminator_shape():     # unittests/test_resnet.py:50-53
    assert simple_discriminator(C.FIX_SIMPLE)(np.ones((1, 128, 128, 3))).numpy().shape == (1, 16, 16, 1)


def test_reflection_padding_golden():      # unittests/test_resnet.py:31-47, bit-exact integer result
    z = np.load(os.path.join(GOLD, "reflect_pad_reference.npz"))
    actual = ReflectionPadding2D()(z["x"]).numpy()
    assert actual.dtype == z["x"].dtype and np.array_equal(z["expected"], actual)
    x = np.random.RandomState(0).randint(-50, 50, size=(2, 5, 7, 3))
    ref = np.pad(x, [(0, 0), (3, 3), (3, 3), (0, 0)], mode="reflect")
    assert np.array_equal(ReflectionPadding2D(padding=(3, 3))(x).numpy(), ref)


# ---- forward parity of every builder -------------------------------------------------------------
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg,size,batch", [
    (C.SMALL_RESNET, 32, 2), (C.SMALL_STRIDED, 32, 2), (C.SMALL_UNET, 40, 2), (C.SMALL_UNET_D, 32, 1),
    (C.SMALL_SIMPLE, 32, 3), (C.FIX_RESNET, 64, 1), (C.UNET_G, 64, 1), (C.UNET_D, 64, 1), (C.SIMPLE_D4, 64, 2),
])
def test_forward_parity(cfg, size, batch, mode):
    m, o = _pair(cfg, mode)
    x = np.random.RandomState(3).uniform(-1, 1, (batch, size, size + 8, 3)).astype(np.float32)   # ragged H != W
    y = m(x).numpy()
    with torch.no_grad():
        ref = o(x).numpy()
    assert y.shape == ref.shape
    assert C.rel_l2(y, ref) <= TOL[mode], C.rel_l2(y, ref)


def test_forward_rejects_bad_inputs():
    m = create_model(C.UNET_G)
    with pytest.raises(ValueError):
        m(np.ones((1, 100, 100, 3)))       # not a multiple of 8
    with pytest.raises(ValueError):
        m(np.ones((1, 64, 64, 4)))


# ---- single-net backward parity through cg_net_forward / cg_net_backward ------------------------------
def _net_grads(m, x, dy):
    import ctypes
    from cyclegan_cat_b200 import _lib
    from cyclegan_cat_b200.runtime import _ptr, _stream_ptr
    lib = _lib.load()
    xd = torch.from_numpy(x).cuda()
    dyd = torch.from_numpy(dy).cuda()
    N, H, W, _ = x.shape
    nbytes = ctypes.c_size_t()
    _lib.check(lib.cg_net_workspace_bytes(m.handle(), N, H, W, 1, ctypes.byref(nbytes)), "ws")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    y = torch.empty(m.out_shape(N, H, W), dtype=torch.float32, device="cuda")
    p = m.device_params()
    g = torch.zeros_like(p)
    dx = torch.empty_like(xd)
    st = _stream_ptr(torch)
    _lib.check(lib.cg_net_forward(m.handle(), _ptr(p), _ptr(xd), _ptr(y), _ptr(ws), ws.numel(), N, H, W, 1, st), "fwd")
    _lib.check(lib.cg_net_backward(m.handle(), _ptr(p), _ptr(dyd), _ptr(dx), _ptr(g), 0, _ptr(ws), ws.numel(), st), "bwd")
    flat = g.cpu().numpy()
    return y.cpu().numpy(), dx.cpu().numpy(), [flat[v.offset:v.offset + v.size].reshape(v.shape) for v in m.trainable_variables]


def _check_grads(got, ref, tol, what=""):
    """relative L2 per variable; variables whose reference gradient is ~0 by construction are
    bounded absolutely by tol * (largest gradient norm in the net)."""
    scale = max(np.linalg.norm(r) for r in ref)
    for i, (g, r) in enumerate(zip(got, ref)):
        nr = np.linalg.norm(r)
        if nr < 1e-6 * scale:
            assert np.linalg.norm(g - r) <= tol * scale, (what, i, np.linalg.norm(g - r), scale)
        else:
            assert C.rel_l2(g, r) <= tol, (what, i, g.shape, C.rel_l2(g, r))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg,size", [(C.SMALL_RESNET, 32), (C.SMALL_STRIDED, 32), (C.SMALL_UNET, 40),
                                      (C.SMALL_UNET_D, 32), (C.SMALL_SIMPLE, 32)])
def test_backward_parity(cfg, size, mode):
    m, o = _pair(cfg, mode)
    rng = np.random.RandomState(5)
    x = rng.uniform(-1, 1, (2, size, size, 3)).astype(np.float32)
    xt = torch.from_numpy(x).requires_grad_(True)
    yo = o.forward(xt)
    dy = rng.normal(0, 1, tuple(yo.shape)).astype(np.float32)
    ref = torch.autograd.grad(yo, [xt] + o.variables, torch.from_numpy(dy))
    y, dx, grads = _net_grads(m, x, dy)
    tol = TOL[mode] * (1 if mode == "fp32" else 2.5)     # bf16 gradients through ~10 layers: 5e-2
    assert C.rel_l2(y, yo.detach().numpy()) <= TOL[mode]
    assert C.rel_l2(dx, ref[0].numpy()) <= tol, C.rel_l2(dx, ref[0].numpy())
    _check_grads(grads, [r.numpy() for r in ref[1:]], tol, cfg["type"])


# ---- the full train step --------------------------------------------------------------------------
def _gan_pair(gen, disc, mode, loss="mse"):
    gan = CycleGan(C.model_config(gen, disc, loss), C.train_config(), mode=mode)
    o = OracleCycleGan(gen, disc, loss=loss)
    for name in ("g_AB", "g_BA", "d_A", "d_B"):
        getattr(gan, name).set_weights([v.detach().numpy() for v in getattr(o, name).variables])
    return gan, o


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("gen,disc,size,batch,loss", [
    (C.SMALL_RESNET, C.SMALL_SIMPLE, 32, 2, "mse"),
    (C.SMALL_UNET, C.SMALL_UNET_D, 32, 1, "mse"),
    (C.SMALL_STRIDED, C.SMALL_SIMPLE, 32, 2, "bce"),
    (C.SMALL_UNET, C.SMALL_SIMPLE, 32, 1, "mae"),
])
def test_train_step_gradients_and_metrics(gen, disc, size, batch, loss, mode):
    gan, o = _gan_pair(gen, disc, mode, loss)
    a, b = synthetic_batch(batch, size)
    ref_m, ref_g, ref_img = o.gradients(a, b)
    m, g = gan.compute_gradients(a, b)
    tol = TOL[mode]
    for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
        assert abs(float(m[k]) - ref_m[k]) <= tol * max(1.0, abs(ref_m[k])), (k, float(m[k]), ref_m[k])
    if mode == "fp32":
        for k in ("dA_acc", "dB_acc"):
            assert abs(float(m[k]) - ref_m[k]) <= 1.0 / (batch * 4), (k, float(m[k]), ref_m[k])
    for name in ("fake_b", "fake_a", "cycled_a", "cycled_b", "same_a", "same_b"):
        assert C.rel_l2(gan.fetch_image(name).numpy(), ref_img[name].numpy()) <= tol, name
    gt = tol if mode == "fp32" else 6e-2
    for net in ("g_AB", "g_BA", "d_A", "d_B"):
        _check_grads(g[net], [r.numpy() for r in ref_g[net]], gt, net)


def test_train_step_c1_three_steps_fp32():
    """Config C1: cycle.yaml U-Net G + simple D [64,128,256], 128x128, batch 1; losses over 3 steps
    and post-step weights."""
    gan, o = _gan_pair(C.UNET_G, C.SIMPLE_D3, "fp32")
    a, b = synthetic_batch(1, 128)
    for step in range(3):
        ref = o.train_step(a, b)
        got = gan.train_step(a, b)
        for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
            assert abs(float(got[k]) - ref[k]) <= 2e-4 * max(1.0, abs(ref[k])), (step, k, float(got[k]), ref[k])
    for name in ("g_AB", "d_A"):
        for v, r in zip(getattr(gan, name).get_weights(), getattr(o, name).variables):
            r = r.detach().numpy()
            if r.ndim == 4:        # kernels: Adam steps of O(lr) on weights of O(0.02)
                assert C.rel_l2(v, r) <= 2e-3, (name, v.shape, C.rel_l2(v, r))
    w = gan.g_AB_optimizer.get_weights()
    assert int(w[0]) == 3 and len(w) == 1 + 2 * len(gan.g_AB.trainable_variables)


def test_frozen_fixture_resnet_small_fp32():
    z = np.load(os.path.join(GOLD, "oracle_resnet_small.npz"))
    gan, _ = _gan_pair(C.SMALL_RESNET, C.SMALL_SIMPLE, "fp32")
    a, b = synthetic_batch(2, 32)
    m, g = gan.compute_gradients(a, b)
    for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
        assert abs(float(m[k]) - float(z[f"metric_{k}"])) <= 1e-4 * max(1.0, abs(float(z[f"metric_{k}"])))
    assert C.rel_l2(gan.fetch_image("fake_b").numpy(), z["fake_b"]) <= 1e-4
    for net in ("g_AB", "g_BA", "d_A", "d_B"):
        for i in (0, 2):
            assert C.rel_l2(g[net][i], z[f"grad_{net}_{i}"]) <= 1e-4, (net, i)


def test_validate_step_matches_and_does_not_train():
    gan, o = _gan_pair(C.SMALL_RESNET, C.SMALL_SIMPLE, "fp32")
    a, b = synthetic_batch(2, 32)
    before = gan.g_AB.get_weights()[0].copy()
    got = gan.validate_step(a, b, training=False)
    ref = o.validate_step(a, b)
    for k in ref:
        assert abs(float(got[k]) - ref[k]) <= 1e-4 * max(1.0, abs(ref[k])), k
    assert np.array_equal(before, gan.g_AB.get_weights()[0])


def test_ragged_last_batch_replans():
    """model.py:197: no drop_remainder -> the last batch is smaller; the trainer re-plans in place."""
    gan, o = _gan_pair(C.SMALL_UNET, C.SMALL_SIMPLE, "fp32")
    a, b = synthetic_batch(3, 32)
    gan.train_step(a, b)
    o.train_step(a, b)
    got = gan.train_step(a[:1], b[:1])
    ref = o.train_step(a[:1], b[:1])
    assert abs(float(got["gAB_loss"]) - ref["gAB_loss"]) <= 2e-4 * abs(ref["gAB_loss"])


def test_linearity_property_of_backward_bf16():
    """Size-independent property at a larger size: backward is linear in dy (scale 2 -> grads x2 exactly in
    structure; checked to bf16 tolerance) for the ResNet generator at 64x64, f=16."""
    m, _ = _pair(C.FIX_RESNET, "bf16")
    rng = np.random.RandomState(1)
    x = rng.uniform(-1, 1, (1, 64, 64, 3)).astype(np.float32)
    dy = rng.normal(0, 1, (1, 64, 64, 3)).astype(np.float32)
    _, _, g1 = _net_grads(m, x, dy)
    _, _, g2 = _net_grads(m, x, 2 * dy)
    big = [i for i, g in enumerate(g1) if g.ndim == 4]
    for i in big:
        assert C.rel_l2(g2[i], 2 * g1[i]) <= 2e-2, i
