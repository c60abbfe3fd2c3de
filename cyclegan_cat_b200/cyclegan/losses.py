"""Loss functions with the reference's names and argument meaning (cyclegan/losses.py:5-81).

Inside `CycleGan.train_step` these are NOT called: the native trainer fuses all of them
(value + gradient seed) into warp-shuffle reduction kernels selected by `LossObj.kind`
(`adv_loss_kernel`, `l1_loss_kernel`, csrc/kernels_elem.cu).  The functions below exist so that
reference-style callers and tests can evaluate a single loss on model outputs.  They are host-side
numpy arithmetic on purpose: device tensors are copied back first, nothing here runs GPU math
outside the library's own kernels.
"""
import numpy as np

from ..ir import LOSS_BY_NAME


def _h(x):
    """Host float64 view of a DeviceTensor / torch tensor / array."""
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        x = x.detach().cpu().numpy() if hasattr(x, "detach") else x.numpy()
    return np.asarray(x, np.float64)


class LossObj:
    """Stands in for the keras loss objects of `get_loss_obj` (losses.py:76-80); SUM_OVER_BATCH_SIZE mean."""

    def __init__(self, name):
        self.name, self.kind = name, LOSS_BY_NAME[name]

    def __call__(self, y_true, y_pred):
        t, p = _h(y_true), _h(y_pred)
        if self.name == "mse":
            return np.float32(((p - t) ** 2).mean())
        if self.name == "mae":
            return np.float32(np.abs(p - t).mean())
        return np.float32((np.maximum(p, 0) - p * t + np.log1p(np.exp(-np.abs(p)))).mean())   # bce from_logits


def calc_cycle_loss(real_image, cycled_image, weight: int = 10):
    return np.float32(weight * np.abs(_h(real_image) - _h(cycled_image)).mean())


def generator_loss(generated, loss_obj: LossObj, weight: float):
    g = _h(generated)
    return np.float32(weight * loss_obj(np.ones_like(g), g))


def identity_loss(real_image, same_image, weight: int = 5):
    return np.float32(weight * np.abs(_h(real_image) - _h(same_image)).mean())


def discriminator_loss(real, generated, loss_obj: LossObj, weight: float):
    r, g = _h(real), _h(generated)
    return np.float32(weight * (loss_obj(np.ones_like(r), r) + loss_obj(np.zeros_like(g), g)))


def get_loss_obj(loss: str) -> LossObj:
    """losses.py:67-81: 'mse' | 'mae' | 'bce' (from_logits), KeyError otherwise."""
    LOSS_OBJ_MAPS = {name: LossObj(name) for name in ("mse", "mae", "bce")}
    return LOSS_OBJ_MAPS[loss]
