"""Layer-by-layer parity harness (test infrastructure): runs the CUDA path, fetches every tensor it stored
(cg_net_fetch_tensor / cg_trainer_fetch_tensor) and replays the oracle with those tensors forced in
(oracle/ir_exec.py).  Gives, for one net call or one whole train step:

  * per-layer forward error  : CUDA output of layer i  vs  oracle op applied to the CUDA path's own input of layer i
  * per-variable gradient error: CUDA gradient  vs  the oracle's exact gradient at the CUDA path's activations
    (ReLU / LeakyReLU masks taken from the stored outputs, so no unit is on different sides of zero)

Gates (BASELINE.json north_star): relative L2 <= 1e-4 in fp32 check mode, <= 2e-2 in bf16 mode; the per-layer forward
gate in bf16 mode is one bf16 ulp (2^-8 = 3.9e-3 -- one storage rounding plus bf16 weights), far inside 2e-2.
Every measured number is appended to gpurun_out/parity/*.jsonl when that directory can be written
(tools/parity_table.py turns them into profiles/r02_parity.md)."""
import json
import os

import numpy as np
import torch

from oracle import models as om
from oracle.ir_exec import IRModel
from oracle.train import OracleCycleGan
from tests import common as C

LAYER_TOL = {"fp32": 1e-4, "bf16": 2.0 ** -8}
GRAD_TOL = {"fp32": 1e-4, "bf16": 2e-2}
OPS = {1: "conv", 2: "convT", 3: "inorm", 4: "act", 5: "rpad", 6: "add", 7: "concat", 8: "avgpool", 9: "upsample",
       10: "bnorm", 11: "dropout"}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, payload):
    """Append one measurement record (never fails the test)."""
    try:
        d = os.path.join(ROOT, "gpurun_out", "parity")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity.jsonl"), "a") as fh:
            fh.write(json.dumps(dict(name=name, **payload)) + "\n")
    except OSError:
        pass


def ir_builder(mode="bf16"):
    """builder(config, dtype) -> IRModel over the graph the PRODUCT builder emits for that config."""
    from cyclegan_cat_b200.cyclegan.model import create_model

    def build(cfg, dtype):
        return IRModel(create_model(cfg, mode=mode).graph, dtype)
    return build


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def to_force(tensors, sl=None):
    """numpy dict -> torch dict (float32 storage; IRModel widens on use); optional batch slice."""
    return {t: torch.from_numpy(v if sl is None else np.ascontiguousarray(v[sl])) for t, v in tensors.items()}


def layer_errors(graph, cuda_tensors, oracle_record, sl=None):
    """[(tensor id, op name, rel L2)] for every tensor the CUDA path stored."""
    out = []
    for t in sorted(cuda_tensors):
        if t == 0 or t not in oracle_record:
            continue
        got = cuda_tensors[t] if sl is None else cuda_tensors[t][sl]
        out.append((t, OPS[graph.layers[t - 1].op], C.rel_l2(got, oracle_record[t].numpy())))
    return out


def grad_errors(got, ref):
    """Per-variable relative L2; variables whose reference gradient is zero by construction (biases feeding an instance
    norm) or tiny by cancellation are measured against 2 % of the net's largest gradient norm (SURVEY 7 'Hard parts')."""
    ref = [np.asarray(r, np.float64) for r in ref]
    scale = max(np.linalg.norm(r) for r in ref)
    return [float(np.linalg.norm(np.asarray(g, np.float64) - r) / max(np.linalg.norm(r), 0.02 * scale))
            for g, r in zip(got, ref)]


def check_single_net(model, oracle, x, dy, mode, net_grads, name):
    """model: product Model; oracle: IRModel with the same weights; net_grads(model, x, dy) -> (y, dx, grads) through
    cg_net_forward(need_backward=1) / cg_net_backward.  Returns the measurement dict after asserting the gates."""
    y, dx, grads = net_grads(model, x, dy)
    tens = model.intermediates()
    rec = {}
    xin = torch.from_numpy(tens[0]).to(oracle.dtype).requires_grad_(True)      # the input as the CUDA path stored it
    yo = oracle.forward(xin, force=to_force(tens), record=rec)
    lerr = layer_errors(model.graph, tens, rec)
    dyr = bf16_round(dy) if mode == "bf16" else dy                               # cg_net_backward converts dy once
    ref = torch.autograd.grad(yo, [xin] + oracle.variables, torch.from_numpy(dyr).to(oracle.dtype))
    gerr = grad_errors(grads, [r.numpy() for r in ref[1:]])
    dxerr = C.rel_l2(dx, ref[0].numpy())
    out = dict(mode=mode, layers=[(t, op, e) for t, op, e in lerr], grads=gerr, dx=dxerr,
               shapes=[list(g.shape) for g in grads])
    record(name, out)
    worst = max(lerr, key=lambda r: r[2])
    assert worst[2] <= LAYER_TOL[mode], (name, "layer", worst)
    assert dxerr <= GRAD_TOL[mode], (name, "dx", dxerr)
    assert max(gerr) <= GRAD_TOL[mode], (name, "grad", int(np.argmax(gerr)), max(gerr))
    return out


# which CUDA call (index into CycleGan.CALLS) and which half of its batch serves each Keras call of model.py:93-106
CALL_OF = {"fake_b": (0, 0), "same_b": (0, 1), "fake_a": (1, 0), "same_a": (1, 1), "cycled_a": (2, None),
           "cycled_b": (3, None), "disc_real_a": (4, 0), "disc_fake_a": (4, 1), "disc_real_b": (5, 0),
           "disc_fake_b": (5, 1)}
NET_OF = {"fake_b": "g_AB", "same_b": "g_AB", "cycled_b": "g_AB", "fake_a": "g_BA", "same_a": "g_BA", "cycled_a": "g_BA",
          "disc_real_a": "d_A", "disc_fake_a": "d_A", "disc_real_b": "d_B", "disc_fake_b": "d_B"}


def gan_pair(gen, disc, mode, loss="mse", dtype=torch.float64):
    """Product CycleGan + the layer-by-layer oracle (IRModels) with identical weights."""
    from cyclegan_cat_b200.cyclegan.model import CycleGan
    gan = CycleGan(C.model_config(gen, disc, loss), C.train_config(), mode=mode)
    o = OracleCycleGan(gen, disc, loss=loss, dtype=dtype, builder=ir_builder(mode),
                       seed_storage=torch.bfloat16 if mode == "bf16" else None)
    for name in ("g_AB", "g_BA", "d_A", "d_B"):
        getattr(gan, name).set_weights([v.detach().numpy() for v in getattr(o, name).variables])
    return gan, o


def check_train_step(gan, o, a, b, mode, name, apply=False):
    """One CycleGan.train_step gradient computation against the teacher-forced oracle: every stored tensor of the six
    model calls, the four losses, and every variable gradient of the four nets.  With apply=True the optimizer step is
    taken as well and the post-step weights are compared (Keras Adam restated in oracle/tf_ops.py)."""
    B = a.shape[0]
    reps = int(os.environ.get("CG_TEST_REPS", "1"))          # diagnostic: >= 3 checks a CUDA-graph replay instead of the eager step
    if os.environ.get("CG_TEST_SIDE_STREAM") == "1":
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(reps):
                metrics, grads = gan.compute_gradients(a, b)
            calls = [gan.call_intermediates(i) for i in range(6)]
        side.synchronize()
    else:
        for _ in range(reps):
            metrics, grads = gan.compute_gradients(a, b)
        calls = [gan.call_intermediates(i) for i in range(6)]
    force = {}
    for key, (ci, half) in CALL_OF.items():
        sl = None if half is None else slice(half * B, (half + 1) * B)
        force[key] = to_force(calls[ci], sl)
    ra, rb = calls[0][0][:B], calls[0][0][B:]              # the real images as the CUDA path stored them
    rec = {}
    ref_m, ref_g, _ = o.gradients(ra, rb, force=force, record=rec)
    lerrs = {}
    for key, (ci, half) in CALL_OF.items():
        sl = None if half is None else slice(half * B, (half + 1) * B)
        graph = getattr(gan, NET_OF[key]).graph
        lerrs[key] = layer_errors(graph, calls[ci], rec[key], sl)
    gerrs = {net: grad_errors(grads[net], [r.numpy() for r in ref_g[net]]) for net in ("g_AB", "g_BA", "d_A", "d_B")}
    merr = {k: abs(float(metrics[k]) - ref_m[k]) / max(1.0, abs(ref_m[k])) for k in ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss")}
    out = dict(mode=mode, batch=int(B), size=int(a.shape[1]),
               layers={k: [(t, op, e) for t, op, e in v] for k, v in lerrs.items()}, grads=gerrs, metrics=merr)
    if apply:
        before = {n: [w.copy() for w in getattr(gan, n).get_weights()] for n in gerrs}
        gan.apply_gradients()
        out["update"] = {}
        for n in gerrs:
            net_o = getattr(o, n)
            opt = getattr(o, n + "_optimizer")
            # (1) the optimizer kernel itself: the oracle's Keras Adam applied to the CUDA gradients
            with torch.no_grad():
                for v, w in zip(net_o.variables, before[n]):
                    v.copy_(torch.from_numpy(w).to(v.dtype))
            opt.apply_gradients([torch.from_numpy(g).to(net_o.dtype) for g in grads[n]], net_o.variables)
            after = getattr(gan, n).get_weights()
            du = [wa - wb for wa, wb in zip(after, before[n])]
            du_ref = [v.detach().numpy() - wb for v, wb in zip(net_o.variables, before[n])]
            num = np.sqrt(sum(float(np.sum((x - y) ** 2)) for x, y in zip(du, du_ref)))
            den = np.sqrt(sum(float(np.sum(y ** 2)) for y in du_ref))
            out["update"][n] = float(num / max(den, 1e-30))
    record(name, out)
    for key, v in lerrs.items():
        worst = max(v, key=lambda r: r[2])
        assert worst[2] <= LAYER_TOL[mode], (name, key, "layer", worst)
    for k, e in merr.items():
        assert e <= (1e-4 if mode == "fp32" else 2e-2), (name, k, e)
    for net, e in gerrs.items():
        assert max(e) <= GRAD_TOL[mode], (name, net, "grad", int(np.argmax(e)), max(e))
    if apply:
        # fp32 weights, fp32 update arithmetic in both: the step lr*m/(sqrt(v)+eps) agrees to float rounding of the weights
        for n, e in out["update"].items():
            assert e <= 1e-3, (name, n, "optimizer update", e)
    return out
