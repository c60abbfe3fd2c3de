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
        self._slots = {}          # unbound use: (model, variable index) -> (m, v) device buffers
        self._iterations = 0

    @property
    def iterations(self):
        if self._binding is not None and self._binding[0]._trainer is not None:
            return self._binding[0]._get_iterations(self._binding[1])
        return self._iterations

    def get_weights(self):
        """Keras order `[iterations, <first slot of every variable>, <second slot ...>]` (what model.py:314-315 saves)."""
        if self._binding is not None and self._binding[0]._trainer is not None:
            return self._binding[0]._optimizer_get_weights(self._binding[1])
        if not self._slots:
            return []               # Keras: no weights before the first apply_gradients
        import numpy as np
        out = [np.int64(self._iterations)]
        per_slot = {"m": 0, "v": 1, "rms": 1}
        for name in self.slots:
            out += [pair[per_slot[name]].cpu().numpy() for pair in self._slots.values()]
        return out

    def set_weights(self, weights):
        if self._binding is None:
            raise RuntimeError("optimizer is not bound to a CycleGan yet")
        self._binding[0]._optimizer_set_weights(self._binding[1], weights)

    def apply_gradients(self, grads_and_vars):
        """keras `optimizer.apply_gradients(zip(grads, variables))` (model.py:149-153; load_optimizer's zero-gradient
        step, model.py:359-362): one fused native launch per variable (`cg_optimizer_apply`), then `iterations += 1`.
        Variables are entries of `model.trainable_variables`; gradients are arrays / tensors of the same shapes.
        An optimizer bound to a `CycleGan` updates that trainer's own slots and iteration counter, so mixing this
        call with `train_step` behaves like Keras; otherwise the slots live in this object."""
        import ctypes

        import numpy as np

        from .. import _lib
        from ..runtime import DeviceTensor, _ptr, _require_cuda, _stream_ptr
        torch = _require_cuda()
        lib = _lib.load()
        cfg = ir.AdamCfg(self.learning_rate, self.beta_1, self.beta_2, self.epsilon, self.kind)
        pairs = [(g, v) for g, v in grads_and_vars if g is not None]
        bound = self._binding is not None and self._binding[0]._trainer is not None
        it = self.iterations if bound else self._iterations
        keep = []
        for g, var in pairs:
            model = var._model
            if var.state:
                raise ValueError("apply_gradients: non-trainable variable")
            p = model.device_params()[var.offset:var.offset + var.size]
            if isinstance(g, DeviceTensor):
                g = g.torch
            gd = g if torch.is_tensor(g) else torch.from_numpy(np.ascontiguousarray(np.asarray(g, np.float32)))
            gd = gd.to(device="cuda", dtype=torch.float32).contiguous().reshape(-1)
            if gd.numel() != var.size:
                raise ValueError(f"gradient of shape {tuple(g.shape)} for variable of shape {var.shape}")
            if bound and model is self._binding[0]._nets()[self._binding[1]]:
                gan, i = self._binding
                m, v = gan._m[i][var.offset:var.offset + var.size], gan._v[i][var.offset:var.offset + var.size]
            else:
                key = (id(model), var.index)
                if key not in self._slots:
                    self._slots[key] = (torch.zeros(var.size, device="cuda"), torch.zeros(var.size, device="cuda"))
                m, v = self._slots[key]
            _lib.check(lib.cg_optimizer_apply(ctypes.byref(cfg), _ptr(p), _ptr(gd), _ptr(m), _ptr(v), var.size, it,
                                              _stream_ptr(torch)), "cg_optimizer_apply")
            keep.append(gd)
        torch.cuda.current_stream().synchronize()       # the uploaded gradients may be freed after this returns
        if bound:
            self._binding[0]._set_iterations(self._binding[1], it + 1)
        else:
            self._iterations = it + 1


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
