"""CPU tests of the host side: reference-compatible builders (KeyError behaviour), IR vs oracle
variable lists, and that the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re
from copy import deepcopy

import numpy as np
import pytest

from cyclegan_cat_b200 import _lib, ir
from cyclegan_cat_b200.cyclegan.model import accuracy, create_model
from cyclegan_cat_b200.cyclegan.optimizers import get_optimizer
from cyclegan_cat_b200.cyclegan.losses import get_loss_obj
from cyclegan_cat_b200.cyclegan.resnet import resnet_generator, simple_discriminator
from cyclegan_cat_b200.cyclegan.unet import strided_unet, unet_generator
from cyclegan_cat_b200.transform.data_load import normalize
from oracle import models as om
from tests import common as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_incomplete_unet_model_config():
    """unittests/test_unet.py:41-56."""
    for field in ['filters', 'kernels', 'expansion', 'normalization', 'dropout', 'output_channels',
                  'final_activation']:
        cfg = deepcopy(C.FIX_UNET)
        del cfg[field]
        with pytest.raises(KeyError):
            unet_generator(cfg)


def test_incomplete_strided_model_config():
    """unittests/test_unet.py:59-72 (strided_unet must NOT require expansion / dropout)."""
    for field in ['filters', 'kernels', 'normalization', 'output_channels', 'final_activation']:
        cfg = deepcopy(C.FIX_UNET)
        del cfg[field]
        with pytest.raises(KeyError):
            strided_unet(cfg)
    cfg = deepcopy(C.FIX_UNET)
    del cfg['expansion'], cfg['dropout']
    strided_unet(cfg)


def test_create_model_dispatch_and_unknown_type():
    """model.py:22-32."""
    assert create_model(C.FIX_SIMPLE).name == "simple_discriminator"
    assert create_model(C.UNET_G).name == "unet_generator"
    with pytest.raises(KeyError):
        create_model(dict(type="nope"))
    with pytest.raises(KeyError):
        resnet_generator({})
    with pytest.raises(KeyError):
        simple_discriminator(dict(filters=[8], kernels=[4]))


@pytest.mark.parametrize("cfg", [C.UNET_G, C.UNET_D, C.SIMPLE_D4, C.RESNET64, C.FIX_UNET, C.SMALL_STRIDED,
                                 C.BN_STRIDED, C.BN_UNET, C.BN_SIMPLE, C.BN_DROP_UNET, C.NONORM_UNET])
def test_variable_lists_match_oracle(cfg):
    """Keras trainable_variables order/shape/initializer agree between the IR builders and the oracle; so do the
    non-trainable ones (BatchNormalization moving statistics)."""
    m = create_model(cfg)
    o = om.create_model(cfg)
    assert m.graph.var_specs() == o.var_specs
    assert [v.shape for v in m.trainable_variables] == [tuple(v.shape) for v in o.variables]
    assert m.n_params == sum(v.numel() for v in o.variables)
    assert [v.shape for v in m.non_trainable_variables] == [tuple(s.shape) for s in o.state]
    assert [v.numpy().tolist() for v in m.non_trainable_variables] == [s.tolist() for s in o.state]     # zeros / ones
    assert m.graph.has_dropout() == (o.n_dropout > 0)
    assert len(m.get_weights()) == len(o.variables) + len(o.state)


def test_optional_paths_plan_natively(built_lib):
    """BatchNormalization / Dropout graphs pass the native planner (pure host code) and report their state size."""
    for cfg, n_bn in ((C.BN_STRIDED, 4), (C.BN_SIMPLE, 2), (C.BN_DROP_UNET, 10), (C.DROP_UNET, 0)):
        m = create_model(cfg)
        n = ctypes.c_int64()
        assert built_lib.cg_net_state_floats(m.handle(), ctypes.byref(n)) == 0
        assert n.value == m.n_state == sum(v.size for v in m.non_trainable_variables)
        assert len(m.non_trainable_variables) == 2 * n_bn
        nbytes = ctypes.c_size_t()
        assert built_lib.cg_net_workspace_bytes(m.handle(), 2, 32, 32, 1, ctypes.byref(nbytes)) == 0 and nbytes.value > 0
    h = ctypes.c_void_p()
    bad = (ir.LayerDesc * 1)(ir.LayerDesc(ir.OP_DROPOUT, 0, -1, 3, 3, 0, 1, 0, 0, 0, 0, 0, 1e-3, 0.2, 0.99, 1.0))   # rate 1
    assert built_lib.cg_net_create(bad, 1, 0, ctypes.byref(h)) == -1 and b"rate" in built_lib.cg_last_error()
    bad = (ir.LayerDesc * 1)(ir.LayerDesc(ir.OP_BNORM, 0, -1, 3, 3, 0, 1, 0, 0, 0, 1, 0, 1e-3, 0.2, 1.5, 0.0))    # momentum
    assert built_lib.cg_net_create(bad, 1, 0, ctypes.byref(h)) == -1 and b"momentum" in built_lib.cg_last_error()


def test_flops_match_survey_tables():
    g = create_model(C.RESNET64).graph
    assert abs(g.flops(256, 256) / 1e9 - 99.103) < 1e-2
    assert abs(create_model(C.SIMPLE_D4).graph.flops(256, 256) / 1e9 - 3.322) < 1e-2
    assert abs(create_model(C.UNET_G).graph.flops(256, 256) / 1e9 - 23.467) < 1e-2
    assert abs(create_model(C.UNET_D).graph.flops(256, 256) / 1e9 - 18.432) < 1e-2


def test_optimizer_and_loss_factories():
    """optimizers.py:5-24, losses.py:67-81."""
    o = get_optimizer(dict(name="adam", learning_rate=2e-4, beta_1=0.5))
    assert (o.learning_rate, o.beta_1, o.beta_2, o.epsilon) == (2e-4, 0.5, 0.999, 1e-7)
    r = get_optimizer(dict(name="rmsprop", learning_rate=1e-3))
    assert (r.kind, r.rho, r.beta_2, r.epsilon, r.slots) == (ir.OPT_RMSPROP, 0.9, 0.9, 1e-7, ("rms",))
    s = get_optimizer(dict(name="sgd", learning_rate=1e-2))
    assert (s.kind, s.learning_rate, s.slots) == (ir.OPT_SGD, 1e-2, ())
    a = get_optimizer(dict(name="adabelief", learning_rate=1e-3))
    assert (a.kind, a.beta_1, a.beta_2, a.epsilon, a.slots) == (ir.OPT_ADABELIEF, 0.9, 0.999, 1e-14, ("m", "v"))
    with pytest.raises(ValueError):
        get_optimizer(dict(name="nadam", learning_rate=1e-3))
    with pytest.raises(KeyError):
        get_optimizer(dict(name="adam", learning_rate=1e-3))          # beta_1 is mandatory for adam
    assert get_loss_obj("mse").kind == 0 and get_loss_obj("bce").kind == 2
    with pytest.raises(KeyError):
        get_loss_obj("hinge")


def test_normalize_and_accuracy():
    """data_load.py:31-34, model.py:35-54."""
    x = np.array([0, 127.5, 255], np.uint8)
    assert np.allclose(normalize(x), [-1.0, 127 / 127.5 - 1, 1.0])
    assert normalize(x).dtype == np.float32
    assert accuracy(np.array([0.9, 0.2]), np.array([0.1, 0.7])) == np.float32(0.5)


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "cyclegan_b200.h")).read()
    declared = set(re.findall(r"^(?:int|void|const char\*)\s+(cg_[a-z0-9_]+)\s*\(", header, re.M))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(built_lib, name), name
    assert built_lib.cg_version() >= 100


def test_struct_layouts_match_header(built_lib):
    assert ctypes.sizeof(ir.LayerDesc) == 16 * 4
    assert ctypes.sizeof(ir.VarInfo) == 40
    assert ctypes.sizeof(ir.TrainCfg) == 5 * 4 + 4 * 20
    lib = _lib.load()
    for which, mirror in enumerate((ir.LayerDesc, ir.VarInfo, ir.TrainCfg, ir.AdamCfg)):
        assert lib.cg_abi_sizeof(which) == ctypes.sizeof(mirror), mirror


def test_native_planner_agrees_with_host(built_lib):
    """cg_net_create / var_info / out_shape are pure host planning: they run without a GPU."""
    for cfg, out in ((C.RESNET64, (2, 64, 96, 3)), (C.SIMPLE_D4, (2, 4, 6, 1)), (C.UNET_G, (2, 64, 96, 3)),
                     (C.SMALL_STRIDED, (2, 64, 96, 3))):
        m = create_model(cfg)
        h = m.handle()
        n = ctypes.c_int()
        assert built_lib.cg_net_var_count(h, ctypes.byref(n)) == 0 and n.value == len(m.trainable_variables)
        for i, v in enumerate(m.trainable_variables):
            info = ir.VarInfo()
            assert built_lib.cg_net_var_info(h, i, ctypes.byref(info)) == 0
            assert tuple(info.shape[:info.ndim]) == v.shape and info.offset == v.offset and info.role == v.role
        assert m.out_shape(2, 64, 96) == out
        nbytes = ctypes.c_size_t()
        assert built_lib.cg_net_workspace_bytes(h, 2, 64, 96, 1, ctypes.byref(nbytes)) == 0 and nbytes.value > 0


def test_trainer_plans_without_a_gpu(built_lib):
    """cg_trainer_create / cg_trainer_workspace_bytes are host planning too.  The step is laid out as five model calls
    (g_AB([a;b]), ONE g_BA call over [fake_b; b; a], g_AB(fake_a), d_A, d_B): the workspace grows linearly with the
    batch, the headline configuration fits a B200 at the benchmark batch (16) and at BASELINE's global batch (64),
    and the fetch probe refuses to run before a step."""
    from cyclegan_cat_b200.cyclegan.model import CycleGan
    gan = CycleGan(C.model_config(C.RESNET64, C.SIMPLE_D4), C.train_config(), mode="bf16")
    cfg = ir.TrainCfg()
    cfg.loss = gan.loss_obj.kind
    for i, o in enumerate(gan._opts()):
        cfg.adam[i] = ir.AdamCfg(o.learning_rate, o.beta_1, o.beta_2, o.epsilon, o.kind)
    tr = ctypes.c_void_p()
    nets = gan._nets()
    assert built_lib.cg_trainer_create(nets[0].handle(), nets[1].handle(), nets[2].handle(), nets[3].handle(),
                                       ctypes.byref(cfg), ctypes.byref(tr)) == 0
    sizes = {}
    for B in (1, 2, 16, 64):
        n = ctypes.c_size_t()
        assert built_lib.cg_trainer_workspace_bytes(tr, B, 256, 256, ctypes.byref(n)) == 0
        sizes[B] = n.value
    assert sizes[2] < 2.2 * sizes[1] and abs(sizes[64] / sizes[16] - 4.0) < 0.05
    assert sizes[16] < 40 * 2 ** 30 and sizes[64] < 150 * 2 ** 30          # of the 180 GB per GPU
    assert built_lib.cg_trainer_workspace_bytes(tr, 1, 250, 250, ctypes.byref(n)) != 0       # 250 % 4 != 0
    shape = (ctypes.c_int * 4)()
    assert built_lib.cg_trainer_fetch_tensor(tr, 1, 0, None, ctypes.byref(shape), None) != 0   # no step yet
    built_lib.cg_trainer_destroy(tr)


def test_native_rejects_bad_graphs_and_shapes(built_lib):
    h = ctypes.c_void_p()
    bad = (ir.LayerDesc * 1)(ir.LayerDesc(ir.OP_CONV, 0, -1, 3, 8, 3, 3, 1, 1, 0, 0, 0, 1e-3, 0.2))   # stride 3
    assert built_lib.cg_net_create(bad, 1, 0, ctypes.byref(h)) == -1
    assert b"geometry" in built_lib.cg_last_error()
    m = create_model(C.UNET_G)
    out = (ctypes.c_int * 4)()
    assert built_lib.cg_net_out_shape(m.handle(), 1, 100, 100, ctypes.byref(out)) == -1    # 100 % 8 != 0
    assert built_lib.cg_net_create(None, 0, 0, ctypes.byref(h)) == -1


def test_no_gpu_means_loud_failure(built_lib):
    """The product path has no CPU fallback: without a device, calling a model raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = create_model(C.FIX_SIMPLE)
    with pytest.raises(_lib.NativeError):
        m(np.ones((1, 32, 32, 3)))


def test_product_package_does_not_import_oracle():
    import subprocess, sys
    code = "import sys; import cyclegan_cat_b200.cyclegan.model, cyclegan_cat_b200.runtime; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_new_entry_points_validate_arguments_without_a_gpu(built_lib):
    """Argument checks of the optional-path / input-pipeline entry points happen before any CUDA call, so they can be
    exercised on a CPU-only box: null pointers, bad geometry, misaligned buffers -> CG_ERR_INVALID (-1) + a message."""
    lib = built_lib
    assert lib.cg_normalize_u8(None, None, 16, None) == -1 and b"null" in lib.cg_last_error()
    assert lib.cg_postprocess_u8(None, None, 16, None) == -1
    assert lib.cg_normalize_u8(ctypes.c_void_p(0x1001), ctypes.c_void_p(0x2000), 16, None) == -1      # src not 4-byte aligned
    assert b"aligned" in lib.cg_last_error()
    assert lib.cg_postprocess_u8(ctypes.c_void_p(0x1008), ctypes.c_void_p(0x2000), 16, None) == -1     # src not 16-byte aligned
    assert lib.cg_normalize_u8(ctypes.c_void_p(0x1000), ctypes.c_void_p(0x2000), 0, None) == 0          # empty input: nothing to do
    p = ctypes.c_void_p(0x1000)
    assert lib.cg_resize_bilinear(None, 1, 8, 8, 3, p, 4, 4, None) == -1
    assert lib.cg_resize_bilinear(p, 1, 0, 8, 3, p, 4, 4, None) == -1 and b"geometry" in lib.cg_last_error()
    assert lib.cg_resize_crop_flip(p, 1, 8, 8, 3, 16, 16, p, 32, 16, None, None, None, None) == -1    # crop larger than the resized image
    assert lib.cg_resize_bilinear(p, 0, 8, 8, 3, p, 4, 4, None) == 0                                      # empty batch
    assert lib.cg_net_state_floats(None, None) == -1
    assert lib.cg_net_set_training(None, 1) == -1 and lib.cg_net_set_seed(None, 1) == -1
    m = create_model(C.BN_SIMPLE)
    assert lib.cg_net_bind_state(m.handle(), None) == -1                                                  # a BatchNormalization net needs its state
    assert lib.cg_net_bind_state(create_model(C.SMALL_SIMPLE).handle(), None) == 0                        # an instance-norm net has none
    assert lib.cg_net_set_training(m.handle(), 1) == 0 and lib.cg_net_set_seed(m.handle(), 2 ** 63 + 5) == 0
    # optimizer kinds are validated when the trainer is created
    cfg = ir.TrainCfg()
    cfg.loss = 0
    for i in range(4):
        cfg.adam[i] = ir.AdamCfg(1e-3, 0.9, 0.999, 1e-7, 7)
    nets = [create_model(C.SMALL_SIMPLE) for _ in range(4)]
    h = ctypes.c_void_p()
    assert lib.cg_trainer_create(*[n.handle() for n in nets], ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"optimizer kind" in lib.cg_last_error()


def test_model_weights_surface_with_batchnorm_state():
    """keras Model.weights / get_weights / set_weights with non-trainable variables, on the host copies (no GPU)."""
    m = create_model(C.BN_SIMPLE)
    nt, ns = len(m.trainable_variables), len(m.non_trainable_variables)
    assert ns == 4 and len(m.weights) == nt + ns
    w = m.get_weights()
    assert [a.shape for a in w[nt:]] == [(8,), (8,), (16,), (16,)]
    assert all(np.all(a == 0) for a in w[nt::2]) and all(np.all(a == 1) for a in w[nt + 1::2])     # moving mean 0, variance 1
    new = [a + 1 for a in w]
    m.set_weights(new)                                      # trainable + non-trainable, Keras order
    assert all(np.array_equal(x, y) for x, y in zip(m.get_weights(), new))
    m.set_weights(w[:nt])                                   # trainable only: the state is left alone
    got = m.get_weights()
    assert all(np.array_equal(x, y) for x, y in zip(got[:nt], w[:nt]))
    assert all(np.array_equal(x, y) for x, y in zip(got[nt:], new[nt:]))
    m.non_trainable_variables[1].assign(np.full((8,), 3.0, np.float32))
    assert np.all(m.non_trainable_variables[1].numpy() == 3.0) and m.non_trainable_variables[1].role == 5
    with pytest.raises(AssertionError):
        m.set_weights(w[:nt - 1])
