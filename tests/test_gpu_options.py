"""GPU parity tests of the optional config paths (BatchNormalization, Dropout, the non-Adam
optimizers) and of the input-pipeline kernels (rank 3), through the C-ABI, against the CPU oracle on identical
weights, state and inputs.  Tolerances: relative L2 <= 1e-4 in fp32 check mode, <= 2e-2 in bf16 mode (BASELINE.json);
integer / byte results bit-exact; gradient gates as in tests/test_gpu_parity.py (_check_grads)."""
import ctypes

import numpy as np
import pytest
import torch

from cyclegan_cat_b200 import _lib
from cyclegan_cat_b200.cyclegan.model import CycleGan, create_model
from cyclegan_cat_b200.runtime import _ptr, _stream_ptr
from cyclegan_cat_b200.transform import data_load as DL
from oracle import models as om, tf_ops as T
from oracle.train import OracleCycleGan, synthetic_batch
from tests import common as C
from tests.test_gpu_parity import _check_grads, _oracle_sensitivity

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "bf16": 2e-2}
NETS = ("g_AB", "g_BA", "d_A", "d_B")


def _pair(cfg, mode, seed=7):
    """Model + fp64 oracle with identical, non-trivial trainable variables AND moving statistics."""
    m = create_model(cfg, mode=mode)
    o = om.create_model(cfg, torch.float64)
    w = om.init_variables(o.var_specs, seed)
    rng = np.random.RandomState(seed + 1)
    w = [a + rng.normal(0, 0.05, a.shape).astype(np.float32) if a.ndim == 1 else a for a in w]
    st = []
    for i, s in enumerate(o.state):
        a = rng.normal(0, 0.1, tuple(s.shape)) if i % 2 == 0 else rng.uniform(0.5, 1.5, tuple(s.shape))
        st.append(a.astype(np.float32))
    m.set_weights(w + st)
    o.load(w)
    for s, a in zip(o.state, st):
        s.copy_(torch.from_numpy(a).double())
    return m, o


def _gan_pair(gen, disc, mode, loss="mse", g_opt=None, d_opt=None):
    gan = CycleGan(C.model_config(gen, disc, loss), C.train_config(g_opt=g_opt, d_opt=d_opt), mode=mode)
    o = OracleCycleGan(gen, disc, loss=loss, g_opt=g_opt, d_opt=d_opt, dtype=torch.float64)
    for i, name in enumerate(NETS):
        getattr(gan, name).set_weights([v.detach().numpy() for v in getattr(o, name).variables])
        getattr(gan, name).set_dropout_seed(100 + i)
        getattr(o, name).drop_seed = 100 + i
    return gan, o


def _state_close(model, oracle_model, tol, what=""):
    for v, s in zip(model.non_trainable_variables, oracle_model.state):
        assert C.rel_l2(v.numpy(), s.numpy()) <= tol, (what, v.index, C.rel_l2(v.numpy(), s.numpy()))


# ---- optimizers (optimizers.py:16-21) ---------------------------------------------------------------------------
@pytest.mark.parametrize("name,lr,start_iter", [("sgd", 1e-2, 0), ("rmsprop", 2e-4, 0), ("adabelief", 2e-4, 0),
                                                ("adabelief", 2e-4, 10)])      # iteration 11+: the rectified branch
def test_optimizer_updates_match_oracle(name, lr, start_iter):
    opt = dict(name=name, learning_rate=lr)
    gan, o = _gan_pair(C.SMALL_STRIDED, C.SMALL_SIMPLE, "fp32", g_opt=opt, d_opt=opt)
    a, b = synthetic_batch(2, 32)
    if start_iter:
        gan.prepare(2, 32, 32)
        for net in NETS:
            oo, go = getattr(o, net + "_optimizer"), getattr(gan, net + "_optimizer")
            oo.iterations = start_iter
            w = go.get_weights()
            go.set_weights([start_iter] + w[1:])
    w0 = {n: [w.copy() for w in getattr(gan, n).get_weights()] for n in NETS}
    for step in range(3):
        ref = o.train_step(a, b)
        got = gan.train_step(a, b)
        lim = 1e-4 if step == 0 else 5e-3
        for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
            assert abs(float(got[k]) - ref[k]) <= lim * max(1.0, abs(ref[k])), (step, k, float(got[k]), ref[k])
    for n in NETS:
        w1 = getattr(gan, n).get_weights()
        for i, (x0, x1, r) in enumerate(zip(w0[n], w1, getattr(o, n).variables)):
            if x0.ndim != 4:
                continue
            d_got, d_ref = x1.astype(np.float64) - x0, r.detach().numpy() - x0
            assert np.linalg.norm(d_ref) > 0
            # the update itself (not the weight, which it changes by ~1 %): wrong coefficients show up as O(1) here
            assert C.rel_l2(d_got, d_ref) <= 0.1, (name, n, i, C.rel_l2(d_got, d_ref))
    nv = len(gan.g_AB.trainable_variables)
    w = gan.g_AB_optimizer.get_weights()
    ow = o.g_AB_optimizer.get_weights()
    assert int(w[0]) == 3 + start_iter == int(ow[0])
    assert len(w) == len(ow) == {"sgd": 1, "rmsprop": 1 + nv, "adabelief": 1 + 2 * nv}[name]
    for x, r in zip(w[1:], ow[1:]):
        if x.ndim == 4:
            assert C.rel_l2(x, r) <= 0.1, (name, "slot", x.shape, C.rel_l2(x, r))


def test_optimizer_state_roundtrip_rmsprop():
    opt = dict(name="rmsprop", learning_rate=2e-4)
    gan, _ = _gan_pair(C.SMALL_STRIDED, C.SMALL_SIMPLE, "fp32", g_opt=opt, d_opt=opt)
    a, b = synthetic_batch(1, 32)
    gan.train_step(a, b)
    w = gan.d_A_optimizer.get_weights()
    gan.d_A_optimizer.set_weights([5] + [x * 2 for x in w[1:]])
    w2 = gan.d_A_optimizer.get_weights()
    assert int(w2[0]) == 5 and all(np.array_equal(x * 2, y) for x, y in zip(w[1:], w2[1:]))


# ---- BatchNormalization ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg,size,batch", [(C.BN_STRIDED, 32, 3), (C.BN_UNET, 32, 2), (C.BN_SIMPLE, 32, 4)])
def test_batchnorm_forward_and_moving_statistics(cfg, size, batch, mode):
    m, o = _pair(cfg, mode)
    x = np.random.RandomState(3).uniform(-1, 1, (batch, size, size + 16, 3)).astype(np.float32)
    tol = TOL[mode] * (2.5 if (mode == "bf16" and cfg is C.BN_UNET) else 1)       # 10 conv + norm stages in bf16
    with torch.no_grad():
        ref_inf = o(x, training=False).numpy()
    assert C.rel_l2(m(x).numpy(), ref_inf) <= tol, ("inference", C.rel_l2(m(x).numpy(), ref_inf))
    _state_close(m, o, 0.0, "inference must not touch the moving statistics")
    with torch.no_grad():
        ref_tr = o(x, training=True).numpy()
    y = m(x, training=True).numpy()
    assert C.rel_l2(y, ref_tr) <= tol, ("training", C.rel_l2(y, ref_tr))
    assert C.rel_l2(ref_tr, ref_inf) > 0.1                          # the two modes really differ on this input
    _state_close(m, o, 1e-5 if mode == "fp32" else 2e-3, "after one training call")
    with torch.no_grad():
        ref2 = o(x, training=False).numpy()
    assert C.rel_l2(m(x).numpy(), ref2) <= tol                      # inference now sees the updated statistics


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg", [C.BN_STRIDED, C.BN_UNET, C.BN_SIMPLE])
def test_batchnorm_backward_parity(cfg, mode):
    m, o = _pair(cfg, mode)
    lib = _lib.load()
    rng = np.random.RandomState(5)
    x = rng.uniform(-1, 1, (3, 32, 32, 3)).astype(np.float32)

    def oracle_grads(xin):
        oo = om.create_model(cfg, torch.float64)
        oo.load([v.detach().numpy() for v in o.variables])
        for s, r in zip(oo.state, o.state):
            s.copy_(r)
        oo.training = True
        xt = torch.from_numpy(np.asarray(xin, np.float64)).requires_grad_(True)
        yo = oo.forward(xt)
        return xt, yo, oo
    xt, yo, oo0 = oracle_grads(x)
    dy = rng.normal(0, 1, tuple(yo.shape)).astype(np.float32)
    ref = torch.autograd.grad(yo, [xt] + oo0.variables, torch.from_numpy(dy).double())
    N, H, W, _c = x.shape
    nbytes = ctypes.c_size_t()
    _lib.check(lib.cg_net_workspace_bytes(m.handle(), N, H, W, 1, ctypes.byref(nbytes)), "ws")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    yd = torch.empty(m.out_shape(N, H, W), dtype=torch.float32, device="cuda")
    p = m.device_params()
    g = torch.zeros_like(p)
    xd, dyd = torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda()
    dx = torch.empty_like(xd)
    st = _stream_ptr(torch)
    _lib.check(lib.cg_net_set_training(m.handle(), 1), "training")
    _lib.check(lib.cg_net_forward(m.handle(), _ptr(p), _ptr(xd), _ptr(yd), _ptr(ws), ws.numel(), N, H, W, 1, st), "fwd")
    _lib.check(lib.cg_net_backward(m.handle(), _ptr(p), _ptr(dyd), _ptr(dx), _ptr(g), 0, _ptr(ws), ws.numel(), st), "bwd")
    flat = g.cpu().numpy()
    grads = [flat[v.offset:v.offset + v.size].reshape(v.shape) for v in m.trainable_variables]
    assert C.rel_l2(yd.cpu().numpy(), yo.detach().numpy()) <= TOL[mode] * (2.5 if mode == "bf16" else 1)
    ref_np = [r.numpy() for r in ref]
    sens = None
    if mode == "fp32":
        def perturbed(eps, seed):
            r2 = np.random.RandomState(seed)
            xp = x.astype(np.float64) * (1 + eps * r2.standard_normal(x.shape))
            xt2, yo2, oo2 = oracle_grads(xp)
            return {"n": [t.numpy() for t in torch.autograd.grad(yo2, [xt2] + oo2.variables, torch.from_numpy(dy).double())]}
        sens = _oracle_sensitivity(perturbed, {"n": ref_np})["n"]
        assert C.rel_l2(dx.cpu().numpy(), ref_np[0]) <= max(1e-4, 2 * sens[0]), C.rel_l2(dx.cpu().numpy(), ref_np[0])
    else:
        assert C.rel_l2(dx.cpu().numpy(), ref_np[0]) <= 0.6
    _check_grads(grads, ref_np[1:], mode, cfg["type"] + "/bn", None if sens is None else sens[1:])
    # an inference-mode forward cannot be back-propagated (the reference only tapes training=True, model.py:138-141)
    _lib.check(lib.cg_net_set_training(m.handle(), 0), "training")
    _lib.check(lib.cg_net_forward(m.handle(), _ptr(p), _ptr(xd), _ptr(yd), _ptr(ws), ws.numel(), N, H, W, 1, st), "fwd")
    assert lib.cg_net_backward(m.handle(), _ptr(p), _ptr(dyd), _ptr(dx), _ptr(g), 0, _ptr(ws), ws.numel(), st) == -4


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_batchnorm_train_steps_match_oracle(mode):
    """strided U-Net G + simple D, both with BatchNormalization, 3 steps (eager, graph capture, graph replay): metrics,
    gradients of step 1, moving statistics after every Keras call of validate_step's order, then an inference step."""
    gan, o = _gan_pair(C.BN_STRIDED, C.BN_SIMPLE, mode)
    a, b = synthetic_batch(2, 32)
    tol = TOL[mode]
    # step 1: gradients
    oc = OracleCycleGan(C.BN_STRIDED, C.BN_SIMPLE, dtype=torch.float64)
    ref_m, ref_g, _ = oc.gradients(a, b)
    m, g = gan.compute_gradients(a, b)
    for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
        assert abs(float(m[k]) - ref_m[k]) <= tol * max(1.0, abs(ref_m[k])), (k, float(m[k]), ref_m[k])
    ref_np = {n: [r.numpy() for r in ref_g[n]] for n in NETS}
    sens = None
    if mode == "fp32":
        def perturbed(eps, seed):
            rng = np.random.RandomState(seed)
            o2 = OracleCycleGan(C.BN_STRIDED, C.BN_SIMPLE, dtype=torch.float64)
            ap = a.astype(np.float64) * (1 + eps * rng.standard_normal(a.shape))
            bp = b.astype(np.float64) * (1 + eps * rng.standard_normal(b.shape))
            return {k: [x.numpy() for x in v] for k, v in o2.gradients(ap, bp)[1].items()}
        sens = _oracle_sensitivity(perturbed, ref_np)
    for n in NETS:
        _check_grads(g[n], ref_np[n], mode, n + "/bn", None if sens is None else sens[n])
    for n in NETS:       # compute_gradients was a training-mode pass: the moving statistics moved exactly as in the oracle
        _state_close(getattr(gan, n), getattr(oc, n), 1e-4 if mode == "fp32" else 5e-2, n)
    # 3 full steps against a fresh oracle that shares the state history: bring `o` to the same point first
    o.gradients(a, b)
    for step in range(3):
        ref = o.train_step(a, b)
        got = gan.train_step(a, b)
        lim = (1e-4 if step == 0 else 5e-3) if mode == "fp32" else 2e-2
        for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
            assert abs(float(got[k]) - ref[k]) <= lim * max(1.0, abs(ref[k])), (step, k, float(got[k]), ref[k])
    for n in NETS:
        _state_close(getattr(gan, n), getattr(o, n), 5e-3 if mode == "fp32" else 5e-2, n + " after 3 steps")
    got = gan.validate_step(a, b, training=False)
    ref = o.validate_step(a, b)
    for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
        assert abs(float(got[k]) - ref[k]) <= (5e-3 if mode == "fp32" else 2e-2) * max(1.0, abs(ref[k])), (k, float(got[k]), ref[k])


def test_batchnorm_checkpoint_keeps_moving_statistics(tmp_path):
    cfg = C.model_config(C.BN_STRIDED, C.BN_SIMPLE)
    cfg.location = str(tmp_path)
    gan = CycleGan(cfg, C.train_config(), mode="fp32")
    a, b = synthetic_batch(2, 32)
    gan.train_step(a, b)
    gan.save_model()
    before = [v.numpy() for v in gan.g_AB.non_trainable_variables]
    assert any(np.abs(x).max() > 0 for x in before[0::2])           # moving means left zero
    assert cfg.new is False                                         # model.py:75-78: the first build cleared it
    gan2 = CycleGan(cfg, C.train_config(), mode="fp32")             # -> load_model()
    for x, v in zip(before, gan2.g_AB.non_trainable_variables):
        assert np.array_equal(x, v.numpy())
    m1, m2 = gan.validate_step(a, b), gan2.validate_step(a, b)
    assert abs(float(m1["gAB_loss"]) - float(m2["gAB_loss"])) <= 1e-6 * abs(float(m1["gAB_loss"]))


# ---- Dropout ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg", [C.DROP_UNET, C.BN_DROP_UNET, C.NONORM_UNET])
def test_dropout_forward_matches_oracle_mask(cfg, mode):
    m, o = _pair(cfg, mode)
    m.set_dropout_seed(4242)
    o.drop_seed = 4242
    x = np.random.RandomState(3).uniform(-1, 1, (2, 32, 48, 3)).astype(np.float32)
    tol = TOL[mode] * (2.5 if mode == "bf16" else 1)
    with torch.no_grad():
        ref_inf = o(x, training=False).numpy()
    assert C.rel_l2(m(x).numpy(), ref_inf) <= tol
    outs = []
    for call in range(2):           # two training calls: two different masks (call counter 0, 1)
        with torch.no_grad():
            ref = o(x, training=True).numpy()
        y = m(x, training=True).numpy()
        assert C.rel_l2(y, ref) <= tol, (call, C.rel_l2(y, ref))
        outs.append(y)
    if cfg is not C.NONORM_UNET:
        assert C.rel_l2(outs[0], outs[1]) > 2 * tol and C.rel_l2(outs[0], ref_inf) > 2 * tol


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_dropout_train_steps_match_oracle(mode):
    """U-Net G with Dropout(0.5) in every double_conv: 3 steps (the third replays the CUDA graph, so the step counter
    must reach the kernels through device memory), every Keras call with its own mask."""
    gan, o = _gan_pair(C.DROP_UNET, C.SMALL_SIMPLE, mode)
    a, b = synthetic_batch(2, 32)
    tol = TOL[mode]
    losses = []
    for step in range(3):
        if step == 0:
            ref_m, ref_g, ref_img = o.gradients(a, b)
            m, g = gan.compute_gradients(a, b)
            for name in ("fake_b", "fake_a", "same_a", "same_b", "cycled_a", "cycled_b"):
                lim = tol * (6 if (mode == "bf16" and name.startswith("cycled")) else 2.5 if mode == "bf16" else 1)
                assert C.rel_l2(gan.fetch_image(name).numpy(), ref_img[name].numpy()) <= lim, name
            ref_np = {n: [r.numpy() for r in ref_g[n]] for n in NETS}
            sens = None
            if mode == "fp32":
                def perturbed(eps, seed):
                    rng = np.random.RandomState(seed)
                    o2 = OracleCycleGan(C.DROP_UNET, C.SMALL_SIMPLE, dtype=torch.float64)
                    for i, n in enumerate(NETS):
                        getattr(o2, n).drop_seed = 100 + i
                    ap = a.astype(np.float64) * (1 + eps * rng.standard_normal(a.shape))
                    bp = b.astype(np.float64) * (1 + eps * rng.standard_normal(b.shape))
                    return {k: [x.numpy() for x in v] for k, v in o2.gradients(ap, bp)[1].items()}
                sens = _oracle_sensitivity(perturbed, ref_np)
            for n in NETS:
                _check_grads(g[n], ref_np[n], mode, n + "/dropout", None if sens is None else sens[n])
            ref = ref_m
            got = m
        else:
            ref = o.train_step(a, b)
            got = gan.train_step(a, b)
        lim = (1e-4 if step == 0 else 5e-3) if mode == "fp32" else 2e-2
        for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
            assert abs(float(got[k]) - ref[k]) <= lim * max(1.0, abs(ref[k])), (step, k, float(got[k]), ref[k])
        losses.append(float(got["gAB_loss"]))
    got = gan.validate_step(a, b)
    ref = o.validate_step(a, b)
    assert abs(float(got["gAB_loss"]) - ref["gAB_loss"]) <= (5e-3 if mode == "fp32" else 2e-2) * abs(ref["gAB_loss"])


def test_dropout_keep_fraction_and_scaling():
    """Statistics of the mask at a larger size: a single Dropout layer on a constant input keeps ~half, scaled by 2."""
    from cyclegan_cat_b200 import ir
    from cyclegan_cat_b200.runtime import Model
    g = ir.Graph()
    g.dropout(g.input, 0.5)
    m = Model(g, name="dropout_only", mode="fp32")
    m.set_dropout_seed(7)
    x = np.ones((4, 128, 128, 3), np.float32)
    y = m(x, training=True).numpy()
    assert set(np.unique(y).tolist()) == {0.0, 2.0}
    assert abs((y == 2.0).mean() - 0.5) < 5e-3
    ref = T.dropout_mask(7, 0, 0, 0, x.size, 0.5).reshape(x.shape) * 2.0
    assert np.array_equal(y, ref)                       # bit-exact: same hash, same element order
    assert np.array_equal(m(x).numpy(), x)              # inference: identity


# ---- input pipeline (data_load.py:20-34, predict.py:26-27) ----------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 1, 1, 3), (2, 5, 7, 3), (16, 256, 256, 3)])
def test_normalize_and_postprocess_bit_exact(shape):
    rng = np.random.RandomState(0)
    u8 = rng.randint(0, 256, size=shape).astype(np.uint8)
    got = DL.normalize_device(u8).numpy()
    assert got.dtype == np.float32 and np.array_equal(got, T.normalize(u8))
    pred = np.tanh(rng.normal(0, 2, size=shape)).astype(np.float32)
    pred.ravel()[:4] = [-1.0, 1.0, 0.0, 0.999][:min(4, pred.size)]
    out = DL.postprocess_prediction(pred)
    assert out.dtype == np.uint8 and np.array_equal(out, T.postprocess_prediction(pred))
    assert np.array_equal(DL.postprocess_prediction(np.full((1, 2, 2, 3), 3.0, np.float32)), np.full((1, 2, 2, 3), 255, np.uint8))


@pytest.mark.parametrize("src,dst", [((2, 13, 17, 3), (26, 34)), ((2, 13, 17, 3), (7, 9)), ((1, 300, 200, 3), (256, 256)),
                                     ((3, 64, 64, 3), (64, 64)), ((8, 286, 286, 3), (128, 128))])
def test_resize_matches_tf_bilinear_restatement(src, dst):
    x = np.random.RandomState(1).uniform(-1, 1, src).astype(np.float32)
    got = DL.resize(x, dst).numpy()
    ref = T.resize_bilinear(x, *dst)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6          # fp32 lerps; the GPU may contract a + (b - a) * t into an fma
    if src[1:3] == dst:
        assert np.array_equal(got, x)               # identity resize is exact


def test_random_jitter_fused_kernel_matches_resize_crop_flip():
    """data_load.py:21-27: resize to size+50, random crop, random flip -- with the draws shared with the oracle."""
    x = np.random.RandomState(2).uniform(-1, 1, (6, 160, 140, 3)).astype(np.float32)
    out, (oy, ox, flip) = DL.random_jitter(x, 128, rng=np.random.RandomState(5), return_draws=True)
    assert flip.min() == 0 and flip.max() == 1 and oy.max() <= 50 and ox.max() <= 50
    ref = T.random_jitter(x, 128, oy, ox, flip)
    assert out.shape == (6, 128, 128, 3)
    assert np.abs(out.numpy() - ref).max() <= 2e-6
    one = DL.random_jitter(x[0], 128, rng=np.random.RandomState(5))
    assert one.shape == (128, 128, 3)
    aug = list(DL.apply_augmentation([x[0], x[1]], 64, rng=np.random.RandomState(1)))
    assert len(aug) == 2 and aug[0].shape == (64, 64, 3) and aug[0].dtype == np.float32


def test_predict_path_with_device_pipeline():
    """predict.py:20-36 end to end on the device: uint8 image -> normalize -> resize -> generator -> uint8."""
    m = create_model(C.FIX_UNET, mode="bf16")
    img = np.random.RandomState(0).randint(0, 256, size=(1, 100, 120, 3)).astype(np.uint8)
    x = DL.resize(DL.normalize_device(img), (64, 64))
    y = m(x)
    out = DL.postprocess_prediction(y)
    assert out.shape == (1, 64, 64, 3) and out.dtype == np.uint8
    assert np.array_equal(out, T.postprocess_prediction(y.numpy()))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_c5_generator_inference_at_512(mode):
    """BASELINE.json configs[4] (C5): cycle.yaml U-Net generator forward at 512x512 (predict.py path), full image size,
    against the fp64 oracle.  bf16 gate: 31 conv layers of bf16 activation storage -- rounding the fp64 oracle's OWN
    activations to bf16 at the points where the CUDA path stores them (no GPU involved) already moves this output by
    5.1e-2 on exactly this input and these weights; the B200 path measures 5.6e-2 (DESIGN.md "bf16 tolerance")."""
    m = create_model(C.UNET_G, mode=mode)
    o = om.create_model(C.UNET_G, torch.float64)
    w = om.init_variables(o.var_specs, 42)
    m.set_weights(w)
    o.load(w)
    u8 = np.random.RandomState(9).randint(0, 256, size=(1, 512, 512, 3)).astype(np.uint8)
    y = m(DL.normalize_device(u8))
    with torch.no_grad():
        ref = o(T.normalize(u8)).numpy()
    assert y.shape == (1, 512, 512, 3)
    assert C.rel_l2(y.numpy(), ref) <= (1e-4 if mode == "fp32" else 8e-2), C.rel_l2(y.numpy(), ref)
    out, want = DL.postprocess_prediction(y), T.postprocess_prediction(ref)
    # uint8 results may differ by one level where (p + 1) * 127.5 sits next to an integer
    if mode == "fp32":
        assert np.abs(out.astype(np.int32) - want.astype(np.int32)).max() <= 1 and (out != want).mean() <= 1e-3


def test_frozen_fixture_optional_paths_fp32():
    """The CUDA path against the committed vectors of tests/golden/oracle_options.npz (BatchNormalization + Dropout U-Net
    G, BatchNormalization simple D, AdaBelief / RMSprop): metrics of two training steps and of an inference step, moving
    statistics, updated weights, a resize sample (the mask bits are compared bit for bit in
    test_dropout_keep_fraction_and_scaling)."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_options.npz"))
    g_opt, d_opt = dict(name="adabelief", learning_rate=2e-4), dict(name="rmsprop", learning_rate=2e-4)
    gan = CycleGan(C.model_config(C.BN_DROP_UNET, C.BN_SIMPLE), C.train_config(g_opt=g_opt, d_opt=d_opt), mode="fp32")
    for i, n in enumerate(NETS):
        getattr(gan, n).initialize(42 + i)              # the oracle's init_variables(seed 42..45) draws the same numbers
        getattr(gan, n).set_dropout_seed(100 + i)
    a, b = synthetic_batch(2, 32)
    for step in range(2):
        got = gan.train_step(a, b)
        for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
            ref = float(z[f"step{step}_{k}"])
            assert abs(float(got[k]) - ref) <= (1e-4 if step == 0 else 5e-3) * max(1.0, abs(ref)), (step, k, float(got[k]), ref)
    got = gan.validate_step(a, b)
    for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
        ref = float(z[f"val_{k}"])
        assert abs(float(got[k]) - ref) <= 5e-3 * max(1.0, abs(ref)), (k, float(got[k]), ref)
    assert C.rel_l2(gan.g_AB.non_trainable_variables[0].numpy(), z["g_AB_moving_mean0"]) <= 5e-3
    assert C.rel_l2(gan.g_AB.non_trainable_variables[1].numpy(), z["g_AB_moving_var0"]) <= 1e-4
    assert C.rel_l2(gan.d_A.non_trainable_variables[3].numpy(), z["d_A_moving_var1"]) <= 1e-4
    assert C.rel_l2(gan.d_A.get_weights()[0], z["d_A_var0"]) <= 1e-3
    assert np.abs(DL.resize(z["resize_in"], (33, 17)).numpy() - z["resize_out"]).max() <= 2e-6
