"""Oracle restatement of ``CycleGan.validate_step`` / ``train_step``.

Follows ``/root/reference/cyclegan/model.py:91-154`` literally: one forward
under a persistent tape, four separate ``tape.gradient`` calls (here four
``torch.autograd.grad`` calls on the same graph), four Keras-Adam updates that
all use the pre-update weights.

Test infrastructure only; PARITY UNPINNED (see ``oracle/__init__.py``).
"""
from typing import Dict

import numpy as np
import torch

from . import tf_ops as T
from .models import create_model, init_variables


class _RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dt):
        ctx.dt = dt
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dt).to(g.dtype), None


def _round_grad(x, dt):
    return _RoundGrad.apply(x, dt)


class OracleCycleGan:
    def __init__(self, gen_config: Dict, disc_config: Dict, loss="mse", loss_weights=None,
                 g_opt=None, d_opt=None, dtype=torch.float32, seeds=(42, 43, 44, 45), builder=None, seed_storage=None):
        """`builder(config, dtype)` defaults to the statement-by-statement builders of oracle/models.py; the layer-by-layer
        tests pass one that returns `oracle.ir_exec.IRModel`s (pinned bit-identical to them, tests/test_oracle.py)."""
        self.dtype = dtype
        # seed_storage=torch.bfloat16: the loss gradients w.r.t. the ten model outputs are rounded to bf16 once, where the
        # CUDA path stores them (the bf16 mode keeps its gradient seeds in the activation dtype); forward values untouched
        self.seed_storage = seed_storage
        builder = builder or create_model
        # model.py:80-89 build_models
        self.g_AB = builder(gen_config, dtype)
        self.g_BA = builder(gen_config, dtype)
        self.d_A = builder(disc_config, dtype)
        self.d_B = builder(disc_config, dtype)
        for net, seed in zip((self.g_AB, self.g_BA, self.d_A, self.d_B), seeds):
            net.load(init_variables(net.var_specs, seed))
        self.loss_obj = T.loss_obj(loss)
        self.loss_weights = loss_weights or dict(cycle=2.0, identity=0.5, generator=1.0, discriminator=0.5)
        g_opt = g_opt or dict(name="adam", learning_rate=2e-4, beta_1=0.5)
        d_opt = d_opt or dict(name="adam", learning_rate=2e-4, beta_1=0.5)
        # model.py:68-71 (optimizers.py:5-24)
        mk = T.get_optimizer
        self.g_AB_optimizer, self.g_BA_optimizer = mk(g_opt), mk(g_opt)
        self.d_A_optimizer, self.d_B_optimizer = mk(d_opt), mk(d_opt)
        self.train_calls = 0         # training-mode steps so far = the dropout counter of the next one

    def nets(self):
        return dict(g_AB=self.g_AB, g_BA=self.g_BA, d_A=self.d_A, d_B=self.d_B)

    def _to(self, x):
        return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(self.dtype)

    def forward_all(self, real_a, real_b, training=False, force=None, record=None):
        """model.py:93-106.  Each model call is one Keras call: its own batch statistics / moving-average update
        (BatchNormalization) and its own dropout mask, keyed by the order of the calls on that model.
        `force` / `record` ({output name: {tensor id: value}}) are handed to IRModel.forward (teacher forcing)."""
        real_a, real_b = self._to(real_a), self._to(real_b)
        calls = {}

        def run(net, x, key):
            net.training = training
            net.call_id = calls.get(id(net), 0)
            net.drop_counter = self.train_calls
            calls[id(net)] = net.call_id + 1
            try:
                if force is None and record is None:
                    return net.forward(x)
                rec = None
                if record is not None:
                    rec = record.setdefault(key, {})
                return net.forward(x, force=None if force is None else force.get(key), record=rec)
            finally:
                net.training = False
        o = {}
        o["fake_b"] = run(self.g_AB, real_a, "fake_b")
        o["cycled_a"] = run(self.g_BA, o["fake_b"], "cycled_a")
        o["fake_a"] = run(self.g_BA, real_b, "fake_a")
        o["cycled_b"] = run(self.g_AB, o["fake_a"], "cycled_b")
        o["same_a"] = run(self.g_BA, real_a, "same_a")
        o["same_b"] = run(self.g_AB, real_b, "same_b")
        o["disc_real_a"] = run(self.d_A, real_a, "disc_real_a")
        o["disc_real_b"] = run(self.d_B, real_b, "disc_real_b")
        o["disc_fake_a"] = run(self.d_A, o["fake_a"], "disc_fake_a")
        o["disc_fake_b"] = run(self.d_B, o["fake_b"], "disc_fake_b")
        if self.seed_storage is not None:
            o = {k: _round_grad(v, self.seed_storage) for k, v in o.items()}
        if training:
            self.train_calls += 1
        return real_a, real_b, o

    def _metrics(self, real_a, real_b, o):
        """model.py:108-134."""
        w = self.loss_weights
        gAB_loss = T.generator_loss(o["disc_fake_b"], self.loss_obj, w["generator"])
        gBA_loss = T.generator_loss(o["disc_fake_a"], self.loss_obj, w["generator"])
        total_cycle = T.calc_cycle_loss(real_a, o["cycled_a"], w["cycle"]) + \
            T.calc_cycle_loss(real_b, o["cycled_b"], w["cycle"])
        total_gAB = gAB_loss + total_cycle + T.identity_loss(real_b, o["same_b"], w["identity"])
        total_gBA = gBA_loss + total_cycle + T.identity_loss(real_a, o["same_a"], w["identity"])
        da_loss = T.discriminator_loss(o["disc_real_a"], o["disc_fake_a"], self.loss_obj, w["discriminator"])
        db_loss = T.discriminator_loss(o["disc_real_b"], o["disc_fake_b"], self.loss_obj, w["discriminator"])
        return dict(gAB_loss=total_gAB, gBA_loss=total_gBA, dA_loss=da_loss, dB_loss=db_loss,
                    dA_acc=T.accuracy(o["disc_real_a"], o["disc_fake_a"]),
                    dB_acc=T.accuracy(o["disc_real_b"], o["disc_fake_b"]))

    def validate_step(self, real_a, real_b, training=False):
        with torch.no_grad():
            ra, rb, o = self.forward_all(real_a, real_b, training)
            return {k: float(v) for k, v in self._metrics(ra, rb, o).items()}

    def gradients(self, real_a, real_b, force=None, record=None):
        """model.py:138-147: returns (metrics, grads dict) without applying them."""
        ra, rb, o = self.forward_all(real_a, real_b, training=True, force=force, record=record)
        metrics = self._metrics(ra, rb, o)
        grads = {}
        for loss_name, net_name in (("gAB_loss", "g_AB"), ("gBA_loss", "g_BA"),
                                    ("dA_loss", "d_A"), ("dB_loss", "d_B")):
            net = getattr(self, net_name)
            grads[net_name] = list(torch.autograd.grad(metrics[loss_name], net.variables, retain_graph=True))
        return {k: float(v.detach()) for k, v in metrics.items()}, grads, {k: v.detach() for k, v in o.items()}

    def train_step(self, real_a, real_b):
        """model.py:136-154."""
        metrics, grads, _ = self.gradients(real_a, real_b)
        for name in ("g_AB", "g_BA", "d_A", "d_B"):    # model.py:149-153
            getattr(self, name + "_optimizer").apply_gradients(grads[name], getattr(self, name).variables)
        return metrics


def synthetic_batch(batch, size, seed_a=1234, seed_b=1235):
    """U(-1,1) float32 NHWC, numpy RandomState so it is reproducible anywhere."""
    a = np.random.RandomState(seed_a).uniform(-1.0, 1.0, size=(batch, size, size, 3)).astype(np.float32)
    b = np.random.RandomState(seed_b).uniform(-1.0, 1.0, size=(batch, size, size, 3)).astype(np.float32)
    return a, b
