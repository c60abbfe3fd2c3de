"""Optimizer factory with the reference's interface (cyclegan/optimizers.py:5-24).

Only Adam -- the optimizer every shipped config uses (configs/training_config.yaml:4-11) --
is built for B200: its update runs as one fused vectorised CUDA kernel over the flat
parameter buffer inside the native trainer (Keras/TF form, epsilon-hat).
"""
from typing import Dict

import numpy as np


class Optimizer:
    pass


class Adam(Optimizer):
    """Host-side description + state accessor.  The slots (m, v) and `iterations` live in
    the native trainer's device buffers once a `CycleGan` binds this optimizer."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self._binding = None      # (CycleGan, slot index) set by CycleGan

    @property
    def iterations(self):
        return self._binding[0]._get_iterations(self._binding[1]) if self._binding else 0

    def get_weights(self):
        """Keras order `[iterations, m_0..m_{n-1}, v_0..v_{n-1}]` (what model.py:314-315 saves)."""
        if self._binding is None:
            return []
        return self._binding[0]._optimizer_get_weights(self._binding[1])

    def set_weights(self, weights):
        if self._binding is None:
            raise RuntimeError("optimizer is not bound to a CycleGan yet")
        self._binding[0]._optimizer_set_weights(self._binding[1], weights)

    def apply_gradients(self, grads_and_vars):
        raise NotImplementedError(
            "standalone apply_gradients is not exposed: CycleGan.train_step applies all four Adam updates in "
            "the native step (cg_trainer_apply_gradients)")


def get_optimizer(optimizer_config: Dict) -> Optimizer:
    learning_rate = optimizer_config["learning_rate"]
    name = optimizer_config["name"]
    if name == "adam":
        optimizer = Adam(learning_rate=learning_rate, beta_1=optimizer_config["beta_1"])
    elif name in ("rmsprop", "sgd", "adabelief"):
        raise NotImplementedError(f"optimizer {name!r} is not built for B200 yet (SURVEY.md 8f rank 4); "
                                  "the shipped configs use adam")
    else:
        raise ValueError(f"Optimizer {name} not found.")
    return optimizer
