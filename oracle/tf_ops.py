"""TF 2.7 / Keras / TFA op semantics restated on torch-CPU (oracle only).

All tensors are NHWC at this level (as in the reference); each op permutes to
NCHW internally for torch.  Kernels keep their TensorFlow layouts:
Conv2D ``(kh, kw, Cin, Cout)``, Conv2DTranspose ``(kh, kw, Cout, Cin)``.

PARITY UNPINNED except ``reflection_pad`` (reference golden vector,
``unittests/test_resnet.py:31-47``): TensorFlow cannot be installed here, so
these follow the published op definitions.
"""
import math

import torch
import torch.nn.functional as F


def same_pad(in_size: int, k: int, s: int):
    """TF ``padding='same'`` amounts (Appendix A.1): extra pixel goes after."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    before = total // 2
    return before, total - before


def conv2d(x, kernel, bias=None, stride=1, padding="same"):
    """``tf.keras.layers.Conv2D`` (call sites: unet.py:25,54,63,121; resnet.py:28,33,40,50,96,103)."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    xn = x.permute(0, 3, 1, 2)
    if padding == "same":
        pt, pb = same_pad(xn.shape[2], kh, stride)
        pl, pr = same_pad(xn.shape[3], kw, stride)
        xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, kernel.permute(3, 2, 0, 1), bias, stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose(x, kernel, bias=None, stride=2):
    """``Conv2DTranspose(padding='same')`` (unet.py:66,76; resnet.py:57), Appendix A.2.

    Full transposed conv, then crop ``[pad_before : pad_before + in*s]`` where
    pad_before is the SAME padding of the forward conv on the output grid.
    """
    kh, kw = kernel.shape[0], kernel.shape[1]
    xn = x.permute(0, 3, 1, 2)
    H, W = xn.shape[2], xn.shape[3]
    # TF kernel (kh,kw,Cout,Cin) -> torch conv_transpose2d weight (Cin, Cout, kh, kw)
    full = F.conv_transpose2d(xn, kernel.permute(3, 2, 0, 1), None, stride=stride)
    pt, _ = same_pad(H * stride, kh, stride)
    pl, _ = same_pad(W * stride, kw, stride)
    y = full[:, :, pt:pt + H * stride, pl:pl + W * stride]
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1)


def instance_norm(x, gamma=None, beta=None, eps=1e-3):
    """TFA ``InstanceNormalization`` (unet.py:30,56,70; resnet.py:29,34,44,51,58,98), Appendix A.3."""
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(1, 2), keepdim=True)
    inv = torch.rsqrt(var + eps)
    if gamma is not None:
        inv = inv * gamma
    y = x * inv + (-mean * inv if beta is None else beta - mean * inv)
    return y


def reflection_pad(x, pad_h: int, pad_w: int):
    """``ReflectionPadding2D.call`` (resnet.py:21-23): tf.pad REFLECT on H and W."""
    if x.dtype in (torch.float32, torch.float64):
        return F.pad(x.permute(0, 3, 1, 2), (pad_w, pad_w, pad_h, pad_h), mode="reflect").permute(0, 2, 3, 1)
    # integer input stays integer (Appendix A.10): index-based mirror
    H, W = x.shape[1], x.shape[2]
    hi = [abs(i) if i < H else 2 * (H - 1) - i for i in range(-pad_h, H + pad_h)]
    wi = [abs(i) if i < W else 2 * (W - 1) - i for i in range(-pad_w, W + pad_w)]
    return x[:, hi][:, :, wi]


def avg_pool2(x):
    """``AveragePooling2D()`` 2x2 / 2 VALID (unet.py:101)."""
    return F.avg_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)


def upsample2(x):
    """``UpSampling2D()`` nearest x2 (unet.py:109)."""
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)


def activation(x, name):
    if name in (None, "linear"):
        return x
    if name == "relu":
        return torch.relu(x)
    if name == "tanh":
        return torch.tanh(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    raise ValueError(f"activation {name!r} not restated")


def leaky_relu(x, alpha=0.2):
    """``LeakyReLU(0.2)`` (resnet.py:101)."""
    return torch.where(x > 0, x, alpha * x)


# --- losses (losses.py), Appendix A.8 -------------------------------------
def loss_obj(name):
    """``get_loss_obj`` (losses.py:67-81): returns f(y_true, y_pred) -> scalar global mean."""
    if name == "mse":
        return lambda t, p: ((p - t) ** 2).mean()
    if name == "mae":
        return lambda t, p: (p - t).abs().mean()
    if name == "bce":
        return lambda t, p: (torch.clamp(p, min=0) - p * t + torch.log1p(torch.exp(-p.abs()))).mean()
    raise KeyError(name)


def calc_cycle_loss(real, cycled, weight):          # losses.py:5-17
    return weight * (real - cycled).abs().mean()


def identity_loss(real, same, weight):              # losses.py:34-46
    return weight * (real - same).abs().mean()


def generator_loss(generated, lobj, weight):        # losses.py:20-31
    return weight * lobj(torch.ones_like(generated), generated)


def discriminator_loss(real, generated, lobj, weight):   # losses.py:49-64
    return weight * (lobj(torch.ones_like(real), real) + lobj(torch.zeros_like(generated), generated))


def accuracy(real, fake):                           # model.py:35-54
    pred = (torch.cat([real, fake], 0) > 0.5).to(torch.float32)
    lab = torch.cat([torch.ones_like(real), torch.zeros_like(fake)], 0).to(torch.float32)
    return (pred == lab).to(torch.float32).mean()


# --- Keras Adam (optimizers.py:14-15), Appendix A.9 ------------------------
class KerasAdam:
    """TF-form Adam: eps added to the un-corrected sqrt(v) (epsilon-hat)."""

    def __init__(self, learning_rate, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self.m = None
        self.v = None

    def apply_gradients(self, grads, variables):
        if self.m is None:
            self.m = [torch.zeros_like(p) for p in variables]
            self.v = [torch.zeros_like(p) for p in variables]
        t = self.iterations + 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        with torch.no_grad():
            for p, g, m, v in zip(variables, grads, self.m, self.v):
                m += (g - m) * (1.0 - self.b1)
                v += (g * g - v) * (1.0 - self.b2)
                p -= lr_t * m / (v.sqrt() + self.eps)
        self.iterations = t

    def get_weights(self):
        """Keras order: [iterations, m_0..m_{n-1}, v_0..v_{n-1}] (model.py:314-315)."""
        return [self.iterations] + [m.numpy().copy() for m in self.m] + [v.numpy().copy() for v in self.v]
