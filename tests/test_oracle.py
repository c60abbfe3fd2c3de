"""CPU tests of the oracle itself: the reference's only known-answer vector, the reference's shape
tests, self-consistency of the TF semantics restated in oracle/tf_ops.py, and the frozen fixtures."""
import math
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


# ---- optional config paths and input pipeline (rank 3): oracle pinned against independent forms ----
def test_batch_norm_against_torch_batch_norm():
    """Appendix A.4: training output uses the biased batch variance, the moving variance the Bessel-corrected one
    (torch's F.batch_norm has the same convention, momentum_torch = 1 - momentum_keras)."""
    from oracle import tf_ops as T
    rng = np.random.RandomState(0)
    x = torch.from_numpy(rng.normal(0.3, 2.0, (3, 5, 4, 6))).double()
    g, b = torch.from_numpy(rng.normal(1, .1, 6)), torch.from_numpy(rng.normal(0, .1, 6))
    state = [torch.zeros(6).double(), torch.ones(6).double()]
    rm, rv = torch.zeros(6).double(), torch.ones(6).double()
    y = T.batch_norm(x, state, g, b, training=True)
    ref = torch.nn.functional.batch_norm(x.permute(0, 3, 1, 2), rm, rv, g, b, True, 0.01, 1e-3).permute(0, 2, 3, 1)
    assert torch.allclose(y, ref, atol=1e-12)
    assert torch.allclose(state[0], rm, atol=1e-12) and torch.allclose(state[1], rv, atol=1e-12)
    y2 = T.batch_norm(x, state, g, b, training=False)
    ref2 = torch.nn.functional.batch_norm(x.permute(0, 3, 1, 2), rm, rv, g, b, False, 0.01, 1e-3).permute(0, 2, 3, 1)
    assert torch.allclose(y2, ref2, atol=1e-12)


def test_dropout_mask_hash_against_python_ints():
    """The numpy (wrapping uint64) restatement of the kernel's splitmix64 mask against arbitrary-precision ints."""
    from oracle import tf_ops as T
    M = (1 << 64) - 1

    def sm(z):
        z = (z + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)
    seed, ctr, call, layer, rate = 0xDEADBEEFCAFE, 7, 2, 3, 0.5
    key = sm(seed ^ ((ctr * 0xD1342543DE82EF95) & M))
    key = sm(key ^ ((call << 32) | layer))
    want = [1.0 if (sm((key + e * 0x9E3779B97F4A7C15) & M) >> 40) / 16777216.0 >= rate else 0.0 for e in range(257)]
    got = T.dropout_mask(seed, ctr, call, layer, 257, rate)
    assert got.tolist() == want
    big = T.dropout_mask(1, 0, 0, 0, 1 << 16, 0.5)
    assert abs(big.mean() - 0.5) < 0.01
    assert not np.array_equal(big, T.dropout_mask(1, 1, 0, 0, 1 << 16, 0.5))      # another step, another mask
    assert not np.array_equal(big, T.dropout_mask(1, 0, 1, 0, 1 << 16, 0.5))      # another call
    assert not np.array_equal(big, T.dropout_mask(1, 0, 0, 1, 1 << 16, 0.5))      # another layer


def test_keras_sgd_rmsprop_and_adabelief_formulas():
    """Appendix A.9 + adabelief_tf's published rule, against a scalar float64 re-derivation."""
    from oracle import tf_ops as T
    g_seq = [0.5, -0.25, 0.125, 0.3, -0.7, 0.2, 0.2]
    # SGD
    p = torch.tensor([1.0], dtype=torch.float64)
    o = T.KerasSGD(0.1)
    for g in g_seq[:2]:
        o.apply_gradients([torch.tensor([g], dtype=torch.float64)], [p])
    assert abs(float(p) - (1.0 - 0.1 * 0.5 + 0.1 * 0.25)) < 1e-15 and o.get_weights() == [2]
    # RMSprop
    p = torch.tensor([1.0], dtype=torch.float64)
    o = T.KerasRMSprop(0.01)
    ref, rms = 1.0, 0.0
    for g in g_seq:
        o.apply_gradients([torch.tensor([g], dtype=torch.float64)], [p])
        rms = 0.9 * rms + 0.1 * g * g
        ref -= 0.01 * g / (math.sqrt(rms) + 1e-7)
    assert abs(float(p) - ref) < 1e-14 and len(o.get_weights()) == 2
    # AdaBelief: steps 1..5 are unrectified (sma_t < 5), later ones rectified
    p = torch.tensor([1.0], dtype=torch.float64)
    o = T.AdaBelief(0.01)
    ref, m, v = 1.0, 0.0, 0.0
    b1, b2, eps = 0.9, 0.999, 1e-14
    rectified = []
    for t, g in enumerate(g_seq, 1):
        o.apply_gradients([torch.tensor([g], dtype=torch.float64)], [p])
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * (g - m) ** 2 + eps
        sma_inf = 2 / (1 - b2) - 1
        sma = sma_inf - 2 * t * b2 ** t / (1 - b2 ** t)
        mc = m / (1 - b1 ** t)
        if sma >= 5:
            r = math.sqrt((sma - 4) / (sma_inf - 4) * (sma - 2) / (sma_inf - 2) * sma_inf / sma)
            ref -= 0.01 * r * mc / (math.sqrt(v / (1 - b2 ** t)) + eps)
        else:
            ref -= 0.01 * mc
        rectified.append(sma >= 5)
        assert abs(float(p) - ref) < 1e-13, t
    assert rectified[0] is False and rectified[-1] is True
    assert len(o.get_weights()) == 3


def test_resize_bilinear_against_torch_interpolate():
    """tf.image.resize bilinear / half-pixel centres == torch's align_corners=False bilinear (no antialias), up and down."""
    from oracle import tf_ops as T
    rng = np.random.RandomState(3)
    x = rng.uniform(0, 255, (2, 13, 17, 3)).astype(np.float32)
    for oh, ow in ((26, 34), (7, 9), (50, 20), (13, 17)):
        ref = torch.nn.functional.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), size=(oh, ow), mode="bilinear",
                                              align_corners=False).permute(0, 2, 3, 1).numpy()
        assert np.abs(T.resize_bilinear(x, oh, ow) - ref).max() <= 1e-3, (oh, ow)
    oy, ox, fl = np.array([3, 0]), np.array([10, 50]), np.array([1, 0])
    j = T.random_jitter(x, 16, oy, ox, fl)
    big = T.resize_bilinear(x, 66, 66)
    assert j.shape == (2, 16, 16, 3)
    assert np.array_equal(j[0], big[0, 3:19, 10:26][:, ::-1]) and np.array_equal(j[1], big[1, 0:16, 50:66])


def test_normalize_and_postprocess_known_answers():
    """data_load.py:31-34, predict.py:26-27."""
    from oracle import tf_ops as T
    assert T.normalize(np.array([0, 255, 51], np.uint8)).tolist() == [-1.0, 1.0, np.float32(51) / np.float32(127.5) - 1]
    assert T.postprocess_prediction(np.array([-1.0, 1.0, 0.0, 0.999], np.float32)).tolist() == [0, 255, 127, 254]


def test_batchnorm_and_dropout_builders_train_step_runs():
    """The optional config paths build and step in the oracle: BN moving statistics move once per Keras call."""
    from oracle.train import OracleCycleGan, synthetic_batch
    gen = dict(C.SMALL_UNET, normalization="batchnorm", dropout=True)
    disc = dict(C.SMALL_SIMPLE, normalization="batchnorm")
    o = OracleCycleGan(gen, disc, g_opt=dict(name="rmsprop", learning_rate=1e-4), d_opt=dict(name="sgd", learning_rate=1e-3))
    a, b = synthetic_batch(2, 16)
    before = [s.clone() for s in o.g_AB.state]
    m0 = o.validate_step(a, b)                      # inference: moving statistics, no dropout, state untouched
    assert all(torch.equal(x, y) for x, y in zip(before, o.g_AB.state))
    m1 = o.train_step(a, b)
    assert not torch.equal(before[0], o.g_AB.state[0])
    assert all(np.isfinite(v) for v in m1.values()) and m0["gAB_loss"] != m1["gAB_loss"]
    assert len(o.d_A.state) == 2 * 2 and len(o.d_A.variables) == 2 * 2 + 2      # BN(center=False, scale=False): no gamma/beta


def test_frozen_fixture_optional_paths():
    """tests/golden/oracle_options.npz (make_golden.py): BatchNormalization + Dropout nets under AdaBelief / RMSprop,
    dropout mask bits, resize / random_jitter samples -- pins the oracle's optional paths against drift."""
    from oracle import tf_ops as T
    z = np.load(os.path.join(GOLD, "oracle_options.npz"))
    o = OracleCycleGan(C.BN_DROP_UNET, C.BN_SIMPLE, g_opt=dict(name="adabelief", learning_rate=2e-4),
                       d_opt=dict(name="rmsprop", learning_rate=2e-4))
    for i, n in enumerate(("g_AB", "g_BA", "d_A", "d_B")):
        getattr(o, n).drop_seed = 100 + i
    a, b = synthetic_batch(2, 32)
    for step in range(2):
        for k, v in o.train_step(a, b).items():
            assert abs(v - float(z[f"step{step}_{k}"])) <= 5e-5 * max(1.0, abs(v)), (step, k, v, float(z[f"step{step}_{k}"]))
    for k, v in o.validate_step(a, b).items():
        assert abs(v - float(z[f"val_{k}"])) <= 5e-4 * max(1.0, abs(v)), (k, v, float(z[f"val_{k}"]))
    assert C.rel_l2(o.g_AB.state[0].numpy(), z["g_AB_moving_mean0"]) < 1e-4
    assert C.rel_l2(o.g_AB.state[1].numpy(), z["g_AB_moving_var0"]) < 1e-5
    assert C.rel_l2(o.d_A.state[3].numpy(), z["d_A_moving_var1"]) < 1e-5
    assert C.rel_l2(o.d_A.variables[0].detach().numpy(), z["d_A_var0"]) < 1e-4
    assert np.array_equal(T.dropout_mask(100, 1, 2, 3, 4096, 0.5).astype(np.uint8), z["dropout_mask"])
    assert np.array_equal(T.resize_bilinear(z["resize_in"], 33, 17), z["resize_out"])
    assert np.array_equal(T.random_jitter(z["resize_in"], 16, np.array([5, 50]), np.array([0, 31]), np.array([1, 0])),
                          z["jitter_out"])


# ---- the layer-by-layer oracle (oracle/ir_exec.py) is the same oracle as the statement-by-statement builders ----------
ALL_CFGS = [C.UNET_G, C.UNET_D, C.SIMPLE_D3, C.FIX_UNET, C.FIX_RESNET, C.FIX_SIMPLE, C.SMALL_RESNET, C.SMALL_STRIDED,
            C.SMALL_UNET, C.SMALL_UNET_D, C.SMALL_SIMPLE, C.BN_STRIDED, C.BN_UNET, C.BN_SIMPLE, C.DROP_UNET,
            C.BN_DROP_UNET, C.NONORM_UNET]


def _ir_pair(cfg, dtype=torch.float64, seed=3):
    from cyclegan_cat_b200.cyclegan.model import create_model     # host-side builder: emits the IR, no GPU touched
    from oracle.ir_exec import IRModel
    g = create_model(cfg).graph
    a, b = models.create_model(cfg, dtype), IRModel(g, dtype)
    w = models.init_variables(a.var_specs, seed)
    rng = np.random.RandomState(seed + 1)
    w = [x + rng.normal(0, 0.05, x.shape).astype(np.float32) if x.ndim == 1 else x for x in w]
    a.load(w)
    b.load(w)
    return a, b


@pytest.mark.parametrize("cfg", ALL_CFGS, ids=lambda c: c["type"])
def test_ir_interpreter_equals_builder_oracle(cfg):
    """Forward, input gradient and every variable gradient are bit-identical: the interpreter issues the same torch
    ops on the same data in the same order as oracle/models.py (which follows unet.py / resnet.py line by line)."""
    a, b = _ir_pair(cfg)
    assert [tuple(v.shape) for v in a.variables] == [tuple(v.shape) for v in b.variables]
    rng = np.random.RandomState(0)
    x = rng.uniform(-1, 1, (2, 32, 48, 3))
    for training in (False, True):
        outs = []
        for m in (a, b):
            m.training, m.call_id, m.drop_counter, m.drop_seed = training, 1, 5, 9
            xt = torch.from_numpy(x).requires_grad_(True)
            y = m.forward(xt)
            dy = torch.from_numpy(np.random.RandomState(1).normal(0, 1, tuple(y.shape)))
            outs.append((y.detach(), torch.autograd.grad(y, [xt] + m.variables, dy, allow_unused=True)))
            m.training = False
        assert torch.equal(outs[0][0], outs[1][0])
        for ga, gb in zip(outs[0][1], outs[1][1]):
            assert (ga is None and gb is None) or torch.equal(ga, gb)
    for sa, sb in zip(a.state, b.state):                 # BatchNormalization moving statistics moved identically
        assert torch.equal(sa, sb)


def test_ir_interpreter_teacher_forcing_and_storage_rounding():
    """force= replaces values but keeps the gradient path; record= holds the oracle's own layer outputs; forcing a tensor
    with its own recorded value changes nothing.  storage=bf16 rounds exactly the tensors the CUDA path stores."""
    _, m = _ir_pair(C.SMALL_RESNET)
    x = torch.from_numpy(np.random.RandomState(0).uniform(-1, 1, (1, 32, 32, 3)))
    rec = {}
    y0 = m.forward(x, record=rec)
    assert len(rec) == len(m.graph.layers)
    y1 = m.forward(x, force={t: v for t, v in rec.items()})
    assert torch.equal(y0, y1)
    # forcing a perturbed tensor: the consumers see the forced VALUE, the producer still receives gradient
    t_mid = 5
    bumped = rec[t_mid] * 1.5
    rec2 = {}
    y2 = m.forward(x, force={t_mid: bumped}, record=rec2)
    assert torch.equal(rec2[t_mid], rec[t_mid]) and not torch.equal(y2, y0)
    g = torch.autograd.grad(y2.sum(), m.variables[0])[0]
    assert float(g.abs().sum()) > 0
    # bf16 storage: every stored tensor is bf16-representable, the norm folded into its ReLU is not rounded on its own
    rec3 = {}
    m.forward(x, record=rec3, storage=torch.bfloat16)
    for t, v in rec3.items():
        if m.stored[t]:
            assert torch.equal(v, v.to(torch.bfloat16).to(v.dtype)), t
    assert not all(m.stored)
