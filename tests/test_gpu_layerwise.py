"""Layer-by-layer GPU parity: every tensor the CUDA path stores and every variable gradient, against the oracle run on
the CUDA path's own activations (tests/layerwise.py, oracle/ir_exec.py).  The gates are BASELINE.json's numbers with
no data-dependent slack: relative L2 <= 1e-4 in fp32 check mode, <= 2e-2 in bf16 mode per variable gradient; per layer
forward <= 1e-4 / one bf16 ulp.  Reference path: cyclegan/model.py:91-154 (validate_step + the four tape.gradient).

The free-running comparisons (no forcing) stay in tests/test_gpu_parity.py; here the full-size headline configuration
(C3) gets its per-layer / per-variable check at the benchmark geometry, plus one batch-16 run."""
import numpy as np
import pytest
import torch

from oracle import models as om
from oracle.ir_exec import IRModel
from oracle.train import OracleCycleGan, synthetic_batch
from tests import common as C
from tests import layerwise as LW
from tests.test_gpu_parity import _net_grads

pytestmark = pytest.mark.gpu

RESNET32 = dict(type="resnet_generator", filters=32)        # 128-channel trunk: every tensor-core layer kind of C3


def _pair(cfg, mode, seed=7, dtype=torch.float64):
    from cyclegan_cat_b200.cyclegan.model import create_model
    m = create_model(cfg, mode=mode)
    o = IRModel(m.graph, dtype)
    w = om.init_variables(o.var_specs, seed)
    rng = np.random.RandomState(seed + 1)
    w = [a + rng.normal(0, 0.05, a.shape).astype(np.float32) if a.ndim == 1 else a for a in w]
    m.set_weights(w)
    o.load(w)
    return m, o


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg,size,batch", [
    (C.SMALL_RESNET, 32, 2), (C.SMALL_STRIDED, 32, 2), (C.SMALL_UNET, 40, 2), (C.SMALL_UNET_D, 32, 1),
    (C.SMALL_SIMPLE, 32, 3), (C.FIX_RESNET, 64, 1), (RESNET32, 64, 2), (C.UNET_G, 64, 2), (C.UNET_D, 64, 1),
    (C.SIMPLE_D4, 64, 2), (C.FIX_UNET, 64, 1), (C.RESNET64, 64, 1),
], ids=lambda v: v["type"] + "_" + str(v.get("filters")).replace(" ", "") if isinstance(v, dict) else str(v))
def test_single_net_layer_by_layer(cfg, size, batch, mode):
    """Every builder, forward and backward, one net call (cg_net_forward / cg_net_backward)."""
    m, o = _pair(cfg, mode)
    rng = np.random.RandomState(5)
    x = rng.uniform(-1, 1, (batch, size, size + 16 * (cfg is not RESNET32), 3)).astype(np.float32)
    shape = m.out_shape(*x.shape[:3])
    dy = rng.normal(0, 1, shape).astype(np.float32)
    LW.check_single_net(m, o, x, dy, mode, _net_grads, f"net/{cfg['type']}/{cfg.get('filters')}/{size}/{mode}")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("gen,disc,size,batch,loss", [
    (C.SMALL_RESNET, C.SMALL_SIMPLE, 32, 2, "mse"),
    (C.SMALL_UNET, C.SMALL_UNET_D, 32, 1, "mse"),
    (C.SMALL_STRIDED, C.SMALL_SIMPLE, 32, 2, "bce"),
    (C.SMALL_UNET, C.SMALL_SIMPLE, 32, 1, "mae"),
    (RESNET32, C.SIMPLE_D3, 64, 2, "mse"),
    (C.UNET_G, C.SIMPLE_D3, 128, 1, "mse"),           # configuration C1
], ids=["resnet8", "unet-unetD", "strided-bce", "unet-mae", "resnet32-tc", "C1"])
def test_train_step_layer_by_layer(gen, disc, size, batch, loss, mode):
    """The whole step: six model calls, four losses, four gradients (model.py:138-147), optimizer update (149-153)."""
    gan, o = LW.gan_pair(gen, disc, mode, loss)
    a, b = synthetic_batch(batch, size)
    LW.check_train_step(gan, o, a, b, mode, f"step/{gen['type']}-{disc['type']}/{size}x{batch}/{loss}/{mode}", apply=True)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_c2_full_size_gradients(mode):
    """configs/cycle.yaml verbatim (C2: U-Net generator + U-Net PatchGAN discriminator) at its full 256x256 size, batch 1:
    every stored tensor, the losses, all variable gradients and the optimizer update.  The geometry the window-form
    kernels (conv_tc_kernel<WIN>, wgradw_tc_kernel) run at in the C2 / C5 benchmarks."""
    dtype = torch.float64 if mode == "fp32" else torch.float32
    gan, o = LW.gan_pair(C.UNET_G, C.UNET_D, mode, dtype=dtype)
    a, b = synthetic_batch(1, 256)
    LW.check_train_step(gan, o, a, b, mode, f"step/C2/256x1/{mode}", apply=True)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_c3_full_size_gradients(mode):
    """The headline configuration at its benchmark geometry (C3: resnet_generator{filters:64} + simple_discriminator
    [64,128,256,512], 256x256; batch 1 so that the CPU oracle finishes in a minute or two): every stored tensor of the
    six model calls, the losses, all 48 + 48 + 10 + 10 variable gradients and the weights after the Adam step.  This is
    the backward the benchmark times: wgrad_tc_kernel, the flat-mode / parity-class data gradients, the fold-mode
    epilogue and the unfolded 7x7 stem / head at 256x256.  bf16 mode is checked against an fp32 oracle (its rounding is
    four orders below the gate), fp32 check mode against fp64."""
    dtype = torch.float64 if mode == "fp32" else torch.float32
    gan, o = LW.gan_pair(C.RESNET64, C.SIMPLE_D4, mode, dtype=dtype)
    a, b = synthetic_batch(1, 256)
    out = LW.check_train_step(gan, o, a, b, mode, f"step/C3/256x1/{mode}", apply=True)
    assert len(out["grads"]["g_AB"]) == 48 and len(out["grads"]["d_A"]) == 10
    assert int(gan.g_AB_optimizer.get_weights()[0]) == 1


def test_c3_batch16_fp32_losses_and_edge_gradients():
    """C3 at the benchmark batch (16 pairs per step) in fp32 check mode, free running against the torch-CPU fp32 oracle:
    the four losses and g_AB's first and last kernel gradient.  (No forcing here -- 16 x the tensors would not fit a
    test -- so the gradient gates carry the ReLU-flip noise of two free-running fp32 implementations.)"""
    from cyclegan_cat_b200.cyclegan.model import CycleGan
    gan = CycleGan(C.model_config(C.RESNET64, C.SIMPLE_D4), C.train_config(), mode="fp32")
    o = OracleCycleGan(C.RESNET64, C.SIMPLE_D4, dtype=torch.float32)
    for name in ("g_AB", "g_BA", "d_A", "d_B"):
        getattr(gan, name).set_weights([v.detach().numpy() for v in getattr(o, name).variables])
    a, b = synthetic_batch(16, 256)
    ref_m, ref_g, _ = o.gradients(a, b)
    m, g = gan.compute_gradients(a, b)
    errs = {}
    for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss"):
        errs[k] = abs(float(m[k]) - ref_m[k]) / max(1.0, abs(ref_m[k]))
        assert errs[k] <= 5e-4, (k, float(m[k]), ref_m[k])
    # The stem kernel's gradient sums every ReLU flip of the whole net: between runs of the SAME build it measures 5e-3 ... 2.1e-2
    # here (the instance-norm statistics are accumulated with atomics, so the flips differ from run to run); the head
    # kernel's gradient has no ReLU downstream and measures 3e-4 ... 1.2e-3.  The tight, deterministic evidence for this
    # geometry is test_c3_full_size_gradients (teacher forced); this test only guards the batch-16 plumbing.
    first, last = 0, len(g["g_AB"]) - 2
    for i, gate in ((first, 6e-2), (last, 2e-2)):
        errs[f"g_AB[{i}]"] = C.rel_l2(g["g_AB"][i], ref_g["g_AB"][i].numpy())
        assert errs[f"g_AB[{i}]"] <= gate, (i, errs)
    LW.record("step/C3/256x16/fp32/free-running", dict(mode="fp32", errors=errs))


@pytest.mark.parametrize("cfg,size", [(C.SMALL_RESNET, 32), (C.FIX_RESNET, 64), (C.SMALL_UNET, 40)])
def test_free_running_bf16_against_storage_emulating_oracle(cfg, size):
    """Reported, loosely gated: the bf16 path free running against (a) the fp64 oracle and (b) the oracle that rounds its
    activations / activation gradients to bf16 at the CUDA path's storage points.  (b) is no closer than (a): two bf16
    pipelines decorrelate to the rounding noise within a few layers, and then disagree on the ReLU units nearest zero
     -- which is why the tight gradient gate above is the teacher-forced one."""
    m, o = _pair(cfg, "bf16")
    rng = np.random.RandomState(5)
    x = rng.uniform(-1, 1, (2, size, size, 3)).astype(np.float32)
    dy = rng.normal(0, 1, m.out_shape(2, size, size)).astype(np.float32)
    y, dx, grads = _net_grads(m, x, dy)
    res = {}
    for label, storage in (("fp64", None), ("bf16-storage", torch.bfloat16)):
        xin = torch.from_numpy(x).double().requires_grad_(True)
        yo = o.forward(xin, storage=storage)
        ref = torch.autograd.grad(yo, [xin] + o.variables, torch.from_numpy(LW.bf16_round(dy)).double())
        e = LW.grad_errors(grads, [r.numpy() for r in ref[1:]])
        res[label] = dict(y=C.rel_l2(y, yo.detach().numpy()), dx=C.rel_l2(dx, ref[0].numpy()), grad_max=max(e),
                          grad_median=float(np.median(e)))
        assert res[label]["y"] <= 2e-2 and res[label]["grad_max"] <= 0.9
    LW.record(f"free/{cfg['type']}/{size}", dict(mode="bf16", **res))


# ---- the tensor-core kernels one by one, against exact fp64 arithmetic on bf16-representable data ----------------------
def _tc_nets():
    from tests.test_gpu_tc import _block_net, _double_conv_net, _stem_head_net, _updown_net
    return [
        # name, graph, (n, h, w), kernels exercised
        ("trunk3x3-256", _block_net(256, 256), (2, 32, 32)),      # conv_tc fwd, flat-mode dgrad (fold epilogue), wgrad_tc
        ("trunk3x3-128x256", _block_net(128, 256), (3, 16, 32)),
        ("down-up-k3", _updown_net(128, 256, 3), (2, 32, 32)),    # stride-2 parity view fwd, parity-class dgrad, convT classes
        ("down-up-k4", _updown_net(64, 128, 4), (2, 32, 64)),     # wgrad_tc transposed roles (Cin = 64)
        ("stem-head-64", _stem_head_net(64), (2, 32, 64)),        # unfolded 7x7 stem / head, stack2 weight gradient
        ("double-conv-k4", _double_conv_net(16, 32, 4), (2, 32, 32)),     # window-form conv (fwd + dgrad) + wgradw_tc
        ("double-conv-k5", _double_conv_net(80, 32, 5), (1, 32, 64)),
        ("double-conv-k3", _double_conv_net(192, 128, 3), (2, 16, 16)),
        ("unet-first-k4", _double_conv_net(3, 16, 4), (2, 64, 128)),      # 3-channel image-side layer: 8-channel re-layout
        ("unet-first-k7", _double_conv_net(3, 16, 7), (1, 128, 256)),     # cycle.yaml discriminator level 0 at full width
        ("double-conv-k4-256", _double_conv_net(16, 16, 4), (1, 64, 256)),   # two 128-pixel tiles per row: one edge each
    ]


@pytest.mark.parametrize("case", range(11))
def test_tc_kernels_against_fp64(case):
    """wgrad_tc_kernel, flat-mode and parity-class data gradients, transposed-conv classes, the stem / head forms and the
    16-channel-group kernels, each inside a two-conv net whose weights and input are exactly representable in bf16 and
    whose every tensor is forced into the fp64 oracle: what is left is fp32 accumulation order and ONE bf16 rounding per
    stored tensor (two for the thin-output forms of the 7x7 layers, whose unfolded intermediate is bf16 as well).
    Gates: every layer output within one bf16 ulp (2^-8 = 3.9e-3), every kernel / input gradient within 1.5 ulp (6e-3);
    the per-channel vectors (gamma, beta, bias: sums over all pixels with heavy cancellation) within 1e-2."""
    from cyclegan_cat_b200.runtime import Model
    name, graph, (n, h, w) = _tc_nets()[case]
    m = Model(graph, name=name, mode="bf16", seed=0)
    o = IRModel(graph, torch.float64)
    rng = np.random.RandomState(2)
    wts = [LW.bf16_round(rng.normal(0, 0.05, v.shape)) for v in m.get_weights()]
    m.set_weights(wts)
    o.load(wts)
    x = LW.bf16_round(rng.uniform(-1, 1, (n, h, w, graph.channels[0])))
    dy = LW.bf16_round(rng.normal(0, 1, m.out_shape(n, h, w)))
    out = LW.check_single_net(m, o, x, dy, "bf16", _net_grads, f"kernel/{name}")
    assert out["dx"] <= 6e-3, (name, "dx", out["dx"])
    for e, shape in zip(out["grads"], out["shapes"]):
        assert e <= (6e-3 if len(shape) == 4 else 1e-2), (name, shape, e)
