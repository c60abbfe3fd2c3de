"""Layer-by-layer oracle: executes a layer graph (the IR the four builders emit, one node per Keras layer of
``/root/reference/cyclegan/unet.py:20-124`` / ``resnet.py:11-105``) with the TF/Keras/TFA op restatements of
``oracle/tf_ops.py``.  ``tests/test_oracle.py`` pins it to the statement-by-statement builders of
``oracle/models.py`` (bit-identical outputs and gradients for every test configuration), so it is the same oracle
with every intermediate tensor addressable by id (tensor 0 = input, layer i -> tensor i + 1).

Two things the closure-style builders cannot do are needed by the GPU parity tests:

* ``force={tensor_id: value}`` -- *teacher forcing*: the value a layer hands to its consumers is replaced by the one the
  CUDA path stored (``cg_net_fetch_tensor``), while the gradient still flows through the oracle's op.  Every layer is
  then checked against the oracle on IDENTICAL inputs (``record`` receives the oracle's own output of each layer), and
  the backward pass is the oracle's exact gradient at the CUDA path's own activations: no ReLU unit can sit on
  different sides of zero in the two implementations, so the per-variable gradient gate measures kernel error only
  (1e-4 in fp32 check mode, 2e-2 in bf16 mode -- BASELINE.json's numbers).
* ``storage=torch.bfloat16`` -- the free-running oracle with its activations (and activation gradients) rounded to
  bf16 wherever the CUDA path stores a tensor.  Reported beside the fp64 comparison (profiles/r02_parity.md); it does
  NOT make a tight gate: two bf16 pipelines that differ by 1e-6 before a rounding decorrelate to the rounding noise
  itself within a handful of layers (each layer maps a relative difference e to sqrt(e * 2^-9)), DESIGN.md section 1.

Test infrastructure only; PARITY UNPINNED (see ``oracle/__init__.py``).
"""
from typing import Dict, List, Optional

import numpy as np
import torch

from . import tf_ops as T

# cg_op / cg_act values of include/cyclegan_b200.h (the IR is plain data: nothing is imported from the product)
(OP_CONV, OP_CONVT, OP_INORM, OP_ACT, OP_RPAD, OP_ADD, OP_CONCAT, OP_AVGPOOL, OP_UPSAMPLE, OP_BNORM,
 OP_DROPOUT) = range(1, 12)
ACT_RELU, ACT_LEAKY, ACT_TANH, ACT_SIGMOID = 1, 2, 3, 4


def _round_fn(dt):
    """y = round_to(dt)(x) in forward, the same rounding on the incoming gradient in backward."""
    class R(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.to(dt).to(x.dtype)

        @staticmethod
        def backward(ctx, g):
            return g.to(dt).to(g.dtype)
    return R.apply


class _PiecewiseLinearAt(torch.autograd.Function):
    """ReLU / LeakyReLU(slope) whose derivative is evaluated at a GIVEN output (the CUDA path's stored activation):
    forward is the plain activation of x, backward multiplies by 1 where that output is > 0 and by `slope` elsewhere.
    With teacher forcing the consumers see the forced output anyway; this makes the derivative belong to the same
    point, so a unit whose pre-activation is within rounding of zero cannot be on in one implementation and off in the
    other (one such unit moves a layer's gradient by ~1/sqrt(units))."""

    @staticmethod
    def forward(ctx, x, out_ref, slope):
        ctx.save_for_backward(out_ref > 0)
        ctx.slope = slope
        return torch.where(x > 0, x, x * slope)

    @staticmethod
    def backward(ctx, g):
        (on,) = ctx.saved_tensors
        return torch.where(on, g, g * ctx.slope), None, None


class IRModel:
    """An ``ir.Graph`` run by the oracle ops.  ``variables`` is in Keras ``trainable_variables`` order."""

    def __init__(self, graph, dtype=torch.float64):
        self.graph, self.dtype = graph, dtype
        self.var_specs = graph.var_specs()
        self.variables: List[torch.Tensor] = [torch.zeros(s, dtype=dtype, requires_grad=True) for s, _ in self.var_specs]
        self.layer_vars, vi = [], 0
        self.state: List[torch.Tensor] = []
        self.layer_state = []
        self.drop_index, nd = [], 0
        for L in graph.layers:
            n = 0
            if L.op in (OP_CONV, OP_CONVT):
                n = 1 + int(bool(L.has_bias))
            elif L.op in (OP_INORM, OP_BNORM) and L.affine:
                n = 2
            self.layer_vars.append(list(range(vi, vi + n)))
            vi += n
            if L.op == OP_BNORM:
                self.layer_state.append(len(self.state))
                self.state += [torch.zeros(L.cin, dtype=dtype), torch.ones(L.cin, dtype=dtype)]
            else:
                self.layer_state.append(-1)
            self.drop_index.append(nd if L.op == OP_DROPOUT else -1)
            nd += int(L.op == OP_DROPOUT)
        assert vi == len(self.variables)
        self.training = False
        self.drop_seed, self.drop_counter, self.call_id = 0, 0, 0
        # the CUDA path folds ReLU / LeakyReLU into the normalisation that feeds only them (csrc/api.cu): the norm's own
        # output is never stored -- the storage emulation must not round there
        ncons = [0] * (len(graph.layers) + 1)
        for L in graph.layers:
            ncons[L.in0] += 1
            if L.op in (OP_ADD, OP_CONCAT):
                ncons[L.in1] += 1
        self.stored = [True] * (len(graph.layers) + 1)
        for i, L in enumerate(graph.layers[:-1]):
            A = graph.layers[i + 1]
            if L.op in (OP_INORM, OP_BNORM) and A.op == OP_ACT and A.in0 == i + 1 and ncons[i + 1] == 1 and \
                    A.act in (ACT_RELU, ACT_LEAKY):
                self.stored[i + 1] = False

    def load(self, arrays):
        assert len(arrays) == len(self.variables)
        with torch.no_grad():
            for v, a in zip(self.variables, arrays):
                assert tuple(v.shape) == tuple(np.shape(a)), (v.shape, np.shape(a))
                v.copy_(torch.as_tensor(np.asarray(a), dtype=self.dtype))

    @property
    def trainable_variables(self):
        return self.variables

    def _layer(self, i, L, t):
        V = [self.variables[j] for j in self.layer_vars[i]]
        x = t[L.in0]
        if L.op == OP_CONV:
            return T.conv2d(x, V[0], V[1] if L.has_bias else None, L.stride, "same" if L.same else "valid")
        if L.op == OP_CONVT:
            return T.conv2d_transpose(x, V[0], V[1] if L.has_bias else None, L.stride)
        if L.op == OP_INORM:
            return T.instance_norm(x, V[0], V[1], eps=L.eps) if L.affine else T.instance_norm(x, eps=L.eps)
        if L.op == OP_BNORM:
            si = self.layer_state[i]
            return T.batch_norm(x, self.state[si:si + 2], V[0] if L.affine else None, V[1] if L.affine else None,
                                training=self.training, eps=L.eps, momentum=L.momentum)
        if L.op == OP_ACT:
            if L.act == ACT_RELU:
                return torch.relu(x)
            if L.act == ACT_LEAKY:
                return T.leaky_relu(x, L.slope)
            return T.activation(x, "tanh" if L.act == ACT_TANH else "sigmoid")
        if L.op == OP_RPAD:
            return T.reflection_pad(x, L.pad, L.pad)
        if L.op == OP_ADD:
            return x + t[L.in1]
        if L.op == OP_CONCAT:
            return torch.cat([x, t[L.in1]], dim=-1)
        if L.op == OP_AVGPOOL:
            return T.avg_pool2(x)
        if L.op == OP_UPSAMPLE:
            return T.upsample2(x)
        if L.op == OP_DROPOUT:
            return T.dropout(x, L.rate, self.training, self.drop_seed, self.drop_counter, self.call_id, self.drop_index[i])
        raise ValueError(L.op)

    def forward(self, x, force: Optional[Dict[int, torch.Tensor]] = None, record: Optional[Dict[int, torch.Tensor]] = None,
                storage=None):
        """x: NHWC tensor.  ``force[t]`` replaces tensor t's VALUE for its consumers (gradient path kept);
        ``record[t]`` receives the oracle's own (pre-forcing) value of every tensor t >= 1."""
        rnd = _round_fn(storage) if storage is not None else None
        t = [x if rnd is None else rnd(x)]
        if force and 0 in force:
            t[0] = t[0] + (force[0].to(self.dtype) - t[0]).detach()
        for i, L in enumerate(self.graph.layers):
            out_ref = None
            if force and L.op == OP_ACT and L.act in (ACT_RELU, ACT_LEAKY):
                if (i + 1) in force:
                    out_ref = force[i + 1]
                else:       # the CUDA path wrote this activation only as the interior of its reflection-padded copy
                    for j in range(i + 1, len(self.graph.layers)):
                        R = self.graph.layers[j]
                        if R.op == OP_RPAD and R.in0 == i + 1 and (j + 1) in force:
                            out_ref = force[j + 1][:, R.pad:-R.pad, R.pad:-R.pad, :]
                            break
            if out_ref is not None:
                y = _PiecewiseLinearAt.apply(t[L.in0], out_ref.to(self.dtype), 0.0 if L.act == ACT_RELU else float(L.slope))
            else:
                y = self._layer(i, L, t)
            if rnd is not None and self.stored[i + 1]:
                y = rnd(y)
            if record is not None:
                record[i + 1] = y.detach()
            if force and (i + 1) in force:
                y = y + (force[i + 1].to(self.dtype) - y).detach()
            t.append(y)
        return t[-1]

    def __call__(self, x, training=False):
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(self.dtype)
        self.training = bool(training)
        try:
            return self.forward(x)
        finally:
            if training:
                self.drop_counter += 1
            self.training = False
