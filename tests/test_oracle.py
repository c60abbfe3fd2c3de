"""CPU tests of the oracle itself: the reference's only known-answer vector, the reference's shape
tests, self-consistency of the TF semantics restated in oracle/tf_ops.py, and the frozen fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import models, tf_ops as T
from oracle.train import OracleCycleGan, synthetic_batch
from tests import common as C

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reflection_padding_reference_golden():
    """unittests/test_resnet.py:31-47 -- the one numeric fixture the reference holds."""
    x = np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2]])[np.newaxis, ..., np.newaxis]
    actual = T.reflection_pad(torch.as_tensor(x), 1, 1).numpy()
    expected = np.array([[1, 1, 1, 1, 1], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1], [2, 2, 2, 2, 2],
                         [1, 1, 1, 1, 1]])[np.newaxis, ..., np.newaxis]
    assert np.array_equal(expected, actual)
    xf = torch.arange(2 * 5 * 4 * 3, dtype=torch.float32).reshape(2, 5, 4, 3)
    assert torch.equal(T.reflection_pad(xf, 2, 2), T.reflection_pad(xf.long(), 2, 2).float())


@pytest.mark.parametrize("builder,cfg,shape", [
    (models.unet_generator, C.FIX_UNET, (1, 128, 128, 3)),      # test_unet.py:27-31
    (models.strided_unet, C.FIX_UNET, (1, 128, 128, 3)),        # test_unet.py:34-38
    (models.resnet_generator, C.FIX_RESNET, (1, 128, 128, 3)),  # test_resnet.py:24-28
    (models.simple_discriminator, C.FIX_SIMPLE, (1, 16, 16, 1)),  # test_resnet.py:50-53
])
def test_reference_shape_tests(builder, cfg, shape):
    m = builder(cfg)
    m.load(models.init_variables(m.var_specs, 0))
    assert tuple(m(np.ones((1, 128, 128, 3))).shape) == shape


def test_param_counts_match_survey():
    n = lambda m: sum(v.numel() for v in m.variables)
    assert n(models.resnet_generator(C.RESNET64)) == 11378179
    assert n(models.simple_discriminator(C.SIMPLE_D4)) == 2757057
    assert n(models.unet_generator(C.UNET_G)) == 1464995
    assert n(models.unet_generator(C.UNET_D)) == 291217


@pytest.mark.parametrize("k", [3, 4, 5, 7])
def test_conv_transpose_is_backprop_of_same_conv(k):
    """Appendix A.2: Conv2DTranspose('same') == conv2d_backprop_input of the SAME conv."""
    torch.manual_seed(k)
    x = torch.randn(2, 6, 5, 2, dtype=torch.float64)
    w = torch.randn(k, k, 3, 2, dtype=torch.float64)
    y = T.conv2d_transpose(x, w, None, 2)
    z = torch.zeros(2, 12, 10, 3, dtype=torch.float64, requires_grad=True)
    gz, = torch.autograd.grad(T.conv2d(z, w, None, 2, "same"), z, x)
    assert (y - gz).abs().max().item() == 0.0


def test_same_padding_table():
    """Appendix A.1 cases."""
    assert T.same_pad(256, 4, 2) == (1, 1)
    assert T.same_pad(256, 3, 2) == (0, 1)
    assert T.same_pad(64, 4, 1) == (1, 2)
    assert T.same_pad(64, 5, 1) == (2, 2)
    assert T.same_pad(64, 7, 1) == (3, 3)
    assert T.same_pad(64, 1, 1) == (0, 0)


def test_instance_norm_definition():
    torch.manual_seed(0)
    x = torch.randn(2, 5, 7, 3, dtype=torch.float64) * 3 + 1
    y = T.instance_norm(x)
    ref = (x - x.mean((1, 2), keepdim=True)) / torch.sqrt(x.var((1, 2), unbiased=False, keepdim=True) + 1e-3)
    assert torch.allclose(y, ref, atol=1e-12)


def test_keras_adam_formula():
    """Appendix A.9 epsilon-hat form differs from torch.optim.Adam; check against a hand loop."""
    p = torch.tensor([1.0, -2.0], dtype=torch.float64)
    g = torch.tensor([0.5, 0.25], dtype=torch.float64)
    opt = T.KerasAdam(2e-4, 0.5)
    opt.apply_gradients([g], [p])
    m = 0.5 * g
    v = 0.001 * g * g
    lr_t = 2e-4 * np.sqrt(1 - 0.999) / (1 - 0.5)
    expect = torch.tensor([1.0, -2.0], dtype=torch.float64) - lr_t * m / (v.sqrt() + 1e-7)
    assert torch.allclose(p, expect, atol=1e-15)
    assert opt.get_weights()[0] == 1 and len(opt.get_weights()) == 3


def test_combined_generator_backward_equals_four_tapes():
    """one backward of L_G = adv_AB + adv_BA + cycle + id_a + id_b equals both tape.gradient calls."""
    o = OracleCycleGan(C.SMALL_RESNET, C.SMALL_SIMPLE, dtype=torch.float64)
    a, b = synthetic_batch(2, 32)
    metrics, grads, _ = o.gradients(a, b)
    ra, rb, out = o.forward_all(a, b)
    w = o.loss_weights
    LG = T.generator_loss(out["disc_fake_b"], o.loss_obj, w["generator"]) + \
        T.generator_loss(out["disc_fake_a"], o.loss_obj, w["generator"]) + \
        T.calc_cycle_loss(ra, out["cycled_a"], w["cycle"]) + T.calc_cycle_loss(rb, out["cycled_b"], w["cycle"]) + \
        T.identity_loss(rb, out["same_b"], w["identity"]) + T.identity_loss(ra, out["same_a"], w["identity"])
    g = torch.autograd.grad(LG, o.g_AB.variables + o.g_BA.variables)
    n = len(o.g_AB.variables)
    for x, y in zip(g[:n], grads["g_AB"]):
        assert torch.allclose(x, y, atol=1e-12, rtol=1e-9)
    for x, y in zip(g[n:], grads["g_BA"]):
        assert torch.allclose(x, y, atol=1e-12, rtol=1e-9)


def test_batched_calls_are_exact_with_instance_norm():
    """g([a; b]) == [g(a); g(b)] because instance-norm statistics are per sample."""
    m = models.resnet_generator(C.SMALL_RESNET, torch.float64)
    m.load(models.init_variables(m.var_specs, 3))
    a, b = synthetic_batch(2, 32)
    both = m(np.concatenate([a, b]))
    assert torch.allclose(both[:2], m(a), atol=1e-12) and torch.allclose(both[2:], m(b), atol=1e-12)


def test_frozen_fixture_train_step():
    """tests/golden/oracle_c1_small.npz was written by tests/golden/make_golden.py from this oracle;
    it pins the oracle against drift (it is NOT a TensorFlow output: parity unpinned)."""
    z = np.load(os.path.join(GOLD, "oracle_c1_small.npz"))
    o = OracleCycleGan(C.SMALL_UNET, C.SMALL_SIMPLE)
    a, b = synthetic_batch(1, 32)
    for step in range(2):
        m = o.train_step(a, b)
        for k, v in m.items():
            assert abs(v - float(z[f"step{step}_{k}"])) <= 2e-5 * max(1.0, abs(v)), (step, k)
    assert C.rel_l2(o.g_AB.variables[0].detach().numpy(), z["g_AB_var0"]) < 1e-5
