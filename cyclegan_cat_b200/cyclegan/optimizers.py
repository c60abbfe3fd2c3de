"""Optimizer factory with the reference's interface (cyclegan/optimizers.py:5-24).

All four optimizers the factory can return are built for B200; each update runs as ONE fused
vectorised CUDA kernel over the flat parameter buffer of a net inside the native trainer
(`cg_trainer_apply_gradients`):

* `Adam`     -- Keras/TF form with epsilon-hat; every shipped config uses it
                (configs/training_config.yaml:4-11);
* `SGD`      -- Keras defaults (momentum 0): p -= lr*g;
* `RMSprop`  -- Keras defaults (rho 0.9, momentum 0, epsilon 1e-7, not centered);
* `AdaBeliefOptimizer` -- adabelief_tf defaults (betas 0.9/0.999, epsilon 1e-14, rectify=True,
                sma_threshold 5, no weight decay, no amsgrad).  adabelief-tf is unpinned in the reference's
                requirements (requirements.txt:13) and absent here: its published update rule is restated.
"""
from typing import Dict

from .. import ir


class Optimizer:
    """Host-side description + state accessor.  The slots and `iterations` live in the native
    trainer's device buffers once a `CycleGan` binds this optimizer."""
    kind = ir.OPT_ADAM
    slots = ("m", "v")            # Keras slot creation order -> get_weights() layout [iterations, slot0..., slot1...]
    beta_1, beta_2, epsilon = 0.0, 0.0, 0.0

    def __init__(self, learning_rate):
        self.learning_rate = learning_rate
        self._binding = None      # (CycleGan, slot index) set by CycleGan

    @property
    def iterations(self):
        return self._binding[0]._get_iterations(self._binding[1]) if self._binding else 0

    def get_weights(self):
        """Keras order `[iterations, <first slot of every variable>, <second slot ...>]` (what model.py:314-315 saves)."""
        if self._binding is None:
            return []
        return self._binding[0]._optimizer_get_weights(self._binding[1])

    def set_weights(self, weights):
        if self._binding is None:
            raise RuntimeError("optimizer is not bound to a CycleGan yet")
        self._binding[0]._optimizer_set_weights(self._binding[1], weights)

    def apply_gradients(self, grads_and_vars):
        raise NotImplementedError(
            "standalone apply_gradients is not exposed: CycleGan.train_step applies all four updates in "
            "the native step (cg_trainer_apply_gradients)")


class Adam(Optimizer):
    kind, slots = ir.OPT_ADAM, ("m", "v")

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        super().__init__(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon


class SGD(Optimizer):
    kind, slots = ir.OPT_SGD, ()

    def __init__(self, learning_rate=0.01):
        super().__init__(learning_rate)


class RMSprop(Optimizer):
    kind, slots = ir.OPT_RMSPROP, ("rms",)

    def __init__(self, learning_rate=0.001, rho=0.9, epsilon=1e-7):
        super().__init__(learning_rate)
        self.rho, self.epsilon = rho, epsilon
        self.beta_2 = rho           # the native config carries rho in the beta_2 field


class AdaBeliefOptimizer(Optimizer):
    kind, slots = ir.OPT_ADABELIEF, ("m", "v")

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-14):
        super().__init__(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon


def get_optimizer(optimizer_config: Dict) -> Optimizer:
    learning_rate = optimizer_config["learning_rate"]
    name = optimizer_config["name"]
    if name == "adam":
        optimizer = Adam(learning_rate=learning_rate, beta_1=optimizer_config["beta_1"])
    elif name == "rmsprop":
        optimizer = RMSprop(learning_rate=learning_rate)
    elif name == "sgd":
        optimizer = SGD(learning_rate=learning_rate)
    elif name == "adabelief":
        optimizer = AdaBeliefOptimizer(learning_rate)
    else:
        raise ValueError(f"Optimizer {name} not found.")
    return optimizer
