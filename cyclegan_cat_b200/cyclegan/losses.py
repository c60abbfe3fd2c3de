"""Loss functions with the reference's names and argument meaning (cyclegan/losses.py:5-81).

Inside `CycleGan.train_step` these are NOT called: the native trainer fuses all of them
(value + gradient seed) into warp-shuffle reduction kernels selected by `LossObj.kind`.
The functions below exist so that reference-style callers and tests can evaluate a
single loss on model outputs; they run on whatever device the tensors live on.
"""
import numpy as np

from ..ir import LOSS_BY_NAME
from ..runtime import DeviceTensor


def _t(x):
    import torch
    if isinstance(x, DeviceTensor):
        return x.torch
    return x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x, np.float32))


class LossObj:
    """Stands in for the keras loss objects of `get_loss_obj` (losses.py:76-80); SUM_OVER_BATCH_SIZE mean."""

    def __init__(self, name):
        self.name, self.kind = name, LOSS_BY_NAME[name]

    def __call__(self, y_true, y_pred):
        import torch
        t, p = _t(y_true), _t(y_pred)
        if self.name == "mse":
            return ((p - t) ** 2).mean()
        if self.name == "mae":
            return (p - t).abs().mean()
        return (torch.clamp(p, min=0) - p * t + torch.log1p(torch.exp(-p.abs()))).mean()   # bce from_logits


def calc_cycle_loss(real_image, cycled_image, weight: int = 10):
    return weight * (_t(real_image) - _t(cycled_image)).abs().mean()


def generator_loss(generated, loss_obj: LossObj, weight: float):
    import torch
    g = _t(generated)
    return weight * loss_obj(torch.ones_like(g), g)


def identity_loss(real_image, same_image, weight: int = 5):
    return weight * (_t(real_image) - _t(same_image)).abs().mean()


def discriminator_loss(real, generated, loss_obj: LossObj, weight: float):
    import torch
    r, g = _t(real), _t(generated)
    return weight * (loss_obj(torch.ones_like(r), r) + loss_obj(torch.zeros_like(g), g))


def get_loss_obj(loss: str) -> LossObj:
    """losses.py:67-81: 'mse' | 'mae' | 'bce' (from_logits), KeyError otherwise."""
    LOSS_OBJ_MAPS = {name: LossObj(name) for name in ("mse", "mae", "bce")}
    return LOSS_OBJ_MAPS[loss]
