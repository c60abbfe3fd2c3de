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


def batch_norm(x, state, gamma=None, beta=None, training=False, eps=1e-3, momentum=0.99):
    """keras ``BatchNormalization`` (unet.py:27-28,57-58,71-72; resnet.py:99-100), Appendix A.4, fused
    implementation (4-D input, axis -1).  ``state`` = [moving_mean, moving_variance] tensors, updated in place
    when training.  Training normalises with the biased batch variance over (N, H, W); the moving variance
    receives the Bessel-corrected one (what ``fused_batch_norm`` returns; Keras leaves the correction in).
    Recalled TF behaviour, unpinned."""
    if training:
        mean = x.mean(dim=(0, 1, 2))
        var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
        n = x.shape[0] * x.shape[1] * x.shape[2]
        with torch.no_grad():
            unbiased = var * (n / (n - 1.0)) if n > 1 else var
            state[0] -= (state[0] - mean.detach()) * (1.0 - momentum)
            state[1] -= (state[1] - unbiased.detach()) * (1.0 - momentum)
    else:
        mean, var = state[0], state[1]
    inv = torch.rsqrt(var + eps)
    if gamma is not None:
        inv = inv * gamma
    return x * inv + (-mean * inv if beta is None else beta - mean * inv)


_M64 = (1 << 64) - 1


def _splitmix64(z):
    """splitmix64 finaliser on numpy uint64 arrays (wrapping arithmetic)."""
    import numpy as np
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def dropout_mask(seed: int, counter: int, call_id: int, layer: int, n: int, rate: float):
    """Keep-mask (float32 0/1, length n) of the B200 library's counter-based dropout: the same hash the CUDA
    kernel evaluates (csrc/kernels_extra.cu dropout_kernel).  TensorFlow's own random stream cannot be
    reproduced, so this is what "the same mask" means for parity; the statistics (keep probability 1-rate,
    inverted scaling) are those of keras ``Dropout`` (unet.py:33-34)."""
    import numpy as np
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed & _M64) ^ (np.uint64(counter) * np.uint64(0xD1342543DE82EF95)))
        key = _splitmix64(key ^ np.uint64(((call_id & 0xFFFFFFFF) << 32) | (layer & 0xFFFFFFFF)))
        e = np.arange(n, dtype=np.uint64)
        h = _splitmix64(key + e * np.uint64(0x9E3779B97F4A7C15))
    u = (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return (u >= np.float32(rate)).astype(np.float32)


def dropout(x, rate, training, seed=0, counter=0, call_id=0, layer=0):
    """keras ``Dropout(rate)``: identity at inference, inverted dropout when training."""
    if not training:
        return x
    keep = torch.from_numpy(dropout_mask(seed, counter, call_id, layer, x.numel(), rate)).to(x.dtype).reshape(x.shape)
    return x * keep * (1.0 / (1.0 - rate))


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


class KerasSGD:
    """``SGD(learning_rate)`` (optimizers.py:19), Keras defaults: momentum 0 -> p -= lr*g; weights = [iterations]."""

    def __init__(self, learning_rate):
        self.lr, self.iterations = learning_rate, 0

    def apply_gradients(self, grads, variables):
        with torch.no_grad():
            for p, g in zip(variables, grads):
                p -= self.lr * g
        self.iterations += 1

    def get_weights(self):
        return [self.iterations]


class KerasRMSprop:
    """``RMSprop(learning_rate)`` (optimizers.py:17), Keras defaults rho .9, momentum 0, eps 1e-7, not centered:
    rms = rho*rms + (1-rho) g^2 ; p -= lr*g/(sqrt(rms)+eps); weights = [iterations, rms...] (Appendix A.9)."""

    def __init__(self, learning_rate, rho=0.9, epsilon=1e-7):
        self.lr, self.rho, self.eps, self.iterations, self.rms = learning_rate, rho, epsilon, 0, None

    def apply_gradients(self, grads, variables):
        if self.rms is None:
            self.rms = [torch.zeros_like(p) for p in variables]
        with torch.no_grad():
            for p, g, r in zip(variables, grads, self.rms):
                r.mul_(self.rho).add_((1.0 - self.rho) * g * g)
                p -= self.lr * g / (r.sqrt() + self.eps)
        self.iterations += 1

    def get_weights(self):
        return [self.iterations] + [r.numpy().copy() for r in self.rms]


class AdaBelief:
    """``AdaBeliefOptimizer(learning_rate)`` (optimizers.py:21) of adabelief-tf (unpinned, requirements.txt:13; the
    package is absent here): published update rule with its defaults beta_1 .9, beta_2 .999, epsilon 1e-14,
    rectify=True, sma_threshold 5, weight_decay 0, amsgrad False, total_steps 0; weights = [iterations, m..., v...]."""

    def __init__(self, learning_rate, beta_1=0.9, beta_2=0.999, epsilon=1e-14, sma_threshold=5.0):
        self.lr, self.b1, self.b2, self.eps, self.thr = learning_rate, beta_1, beta_2, epsilon, sma_threshold
        self.iterations, self.m, self.v = 0, None, None

    def apply_gradients(self, grads, variables):
        if self.m is None:
            self.m = [torch.zeros_like(p) for p in variables]
            self.v = [torch.zeros_like(p) for p in variables]
        t = self.iterations + 1
        b1p, b2p = self.b1 ** t, self.b2 ** t
        sma_inf = 2.0 / (1.0 - self.b2) - 1.0
        sma_t = sma_inf - 2.0 * t * b2p / (1.0 - b2p)
        with torch.no_grad():
            for p, g, m, v in zip(variables, grads, self.m, self.v):
                m.mul_(self.b1).add_((1.0 - self.b1) * g)
                v.mul_(self.b2).add_((1.0 - self.b2) * (g - m) ** 2 + self.eps)
                m_corr = m / (1.0 - b1p)
                if sma_t >= self.thr:
                    r_t = math.sqrt((sma_t - 4.0) / (sma_inf - 4.0) * (sma_t - 2.0) / (sma_inf - 2.0) * sma_inf / sma_t)
                    step = r_t * m_corr / ((v / (1.0 - b2p)).sqrt() + self.eps)
                else:
                    step = m_corr
                p -= self.lr * step
        self.iterations = t

    def get_weights(self):
        return [self.iterations] + [m.numpy().copy() for m in self.m] + [v.numpy().copy() for v in self.v]


def get_optimizer(cfg):
    """``get_optimizer`` (optimizers.py:5-24)."""
    name, lr = cfg["name"], cfg["learning_rate"]
    if name == "adam":
        return KerasAdam(lr, cfg["beta_1"])
    if name == "rmsprop":
        return KerasRMSprop(lr)
    if name == "sgd":
        return KerasSGD(lr)
    if name == "adabelief":
        return AdaBelief(lr)
    raise ValueError(f"Optimizer {name} not found.")


# --- input pipeline (transform/data_load.py:20-34, predict.py:26-27), numpy --------------------------------------
def normalize(x):
    """data_load.py:31-34."""
    import numpy as np
    return np.asarray(x, np.float32) / np.float32(127.5) - np.float32(1.0)


def postprocess_prediction(pred):
    """predict.py:26-27 for a whole batch: np.array((p + 1) * 127.5, np.uint8) (truncation; range clamped)."""
    import numpy as np
    t = (np.asarray(pred, np.float32) + np.float32(1.0)) * np.float32(127.5)
    return np.clip(np.trunc(t), 0, 255).astype(np.uint8)


def resize_bilinear(x, out_h, out_w):
    """``tf.image.resize(x, [out_h, out_w])`` (data_load.py:23,41): bilinear, half_pixel_centers=True, antialias=False,
    float32 arithmetic in TF's order (lerp along x, then along y).  x: numpy NHWC."""
    import numpy as np
    x = np.asarray(x, np.float32)
    N, H, W, C = x.shape

    def coords(out, size):
        scale = np.float32(size) / np.float32(out)
        src = (np.arange(out, dtype=np.float32) + np.float32(0.5)) * scale - np.float32(0.5)
        f = np.floor(src)
        lo = np.maximum(f, 0).astype(np.int64)
        hi = np.minimum(np.ceil(src), size - 1).astype(np.int64)
        return lo, hi, (src - f).astype(np.float32)

    y0, y1, ly = coords(out_h, H)
    x0, x1, lx = coords(out_w, W)
    lx = lx[None, None, :, None]
    ly = ly[None, :, None, None]
    top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * lx
    bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * lx
    return (top + (bot - top) * ly).astype(np.float32)


def random_jitter(x, image_size, oy, ox, flip):
    """data_load.py:21-27 with the random draws made explicit: resize to image_size+50, crop at (oy, ox), mirror."""
    import numpy as np
    big = resize_bilinear(x, image_size + 50, image_size + 50)
    out = np.stack([big[n, oy[n]:oy[n] + image_size, ox[n]:ox[n] + image_size] for n in range(len(big))])
    for n in range(len(out)):
        if flip[n]:
            out[n] = out[n][:, ::-1]
    return out
