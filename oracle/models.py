"""Oracle restatement of the four reference model builders (torch-CPU, NHWC).

Follows ``/root/reference/cyclegan/unet.py:20-124`` and
``/root/reference/cyclegan/resnet.py:11-105`` statement by statement.  Each
builder returns an ``OracleModel`` whose ``variables`` list is in Keras
``trainable_variables`` order (forward-chain order; Conv -> [kernel, bias],
InstanceNormalization -> [gamma, beta]).

Test infrastructure only; PARITY UNPINNED (see ``oracle/__init__.py``).
"""
from typing import Dict, List

import numpy as np
import torch

from . import tf_ops as T


class OracleModel:
    def __init__(self, dtype=torch.float32):
        self.dtype = dtype
        self.variables: List[torch.Tensor] = []
        self.var_specs = []          # (shape, init kind) in creation order
        self._program = []           # list of callables built by the builder
        # keras non_trainable_variables: [moving_mean, moving_variance] per BatchNormalization
        self.state: List[torch.Tensor] = []
        self.training = False        # the `training=` argument of the current call
        self.n_dropout = 0           # dropout layers created so far (their ordinal keys the mask)
        self.drop_seed, self.drop_counter, self.call_id = 0, 0, 0

    # -- variable creation ---------------------------------------------------
    def _var(self, shape, init):
        self.var_specs.append((tuple(shape), init))
        v = torch.zeros(shape, dtype=self.dtype, requires_grad=True)
        self.variables.append(v)
        return len(self.variables) - 1

    def load(self, arrays):
        assert len(arrays) == len(self.variables)
        with torch.no_grad():
            for v, a in zip(self.variables, arrays):
                assert tuple(v.shape) == tuple(a.shape), (v.shape, a.shape)
                v.copy_(torch.as_tensor(np.asarray(a), dtype=self.dtype))

    @property
    def trainable_variables(self):
        return self.variables

    def __call__(self, x, training=False):
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(self.dtype)
        self.training = bool(training)
        try:
            return self.forward(x)
        finally:
            if training:
                self.drop_counter += 1          # standalone calls: one dropout counter tick per training call
            self.training = False


def init_variables(var_specs, seed):
    """Shared deterministic init: N(0,0.02) kernels, zero biases,
    ones/zeros for IN gamma/beta, glorot-uniform for the unet 1x1 head (unet.py:121)."""
    rng = np.random.RandomState(seed)
    out = []
    for shape, kind in var_specs:
        if kind == "normal":
            out.append(rng.normal(0.0, 0.02, size=shape).astype(np.float32))
        elif kind == "zeros":
            out.append(np.zeros(shape, np.float32))
        elif kind == "ones":
            out.append(np.ones(shape, np.float32))
        elif kind == "glorot":
            kh, kw, cin, cout = shape
            lim = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
            out.append(rng.uniform(-lim, lim, size=shape).astype(np.float32))
        else:
            raise ValueError(kind)
    return out


# ---------------------------------------------------------------------------
# layer helpers: each appends variables and returns a closure over their index
# ---------------------------------------------------------------------------
def _conv(m: OracleModel, cin, cout, k, stride, padding, use_bias=True, init="normal"):
    ki = m._var((k, k, cin, cout), init)
    bi = m._var((cout,), "zeros") if use_bias else None
    return lambda x: T.conv2d(x, m.variables[ki], None if bi is None else m.variables[bi], stride, padding)


def _convT(m: OracleModel, cin, cout, k, stride):
    ki = m._var((k, k, cout, cin), "normal")
    bi = m._var((cout,), "zeros")
    return lambda x: T.conv2d_transpose(x, m.variables[ki], m.variables[bi], stride)


def _inorm(m: OracleModel, c, affine):
    if affine:
        gi = m._var((c,), "ones")
        bi = m._var((c,), "zeros")
        return lambda x: T.instance_norm(x, m.variables[gi], m.variables[bi])
    return lambda x: T.instance_norm(x)


def _bnorm(m: OracleModel, c, affine):
    """keras BatchNormalization(): trainable [gamma, beta] when center/scale, state [moving_mean 0, moving_variance 1]."""
    gi = m._var((c,), "ones") if affine else None
    bi = m._var((c,), "zeros") if affine else None
    si = len(m.state)
    m.state += [torch.zeros(c, dtype=m.dtype), torch.ones(c, dtype=m.dtype)]
    return lambda x: T.batch_norm(x, m.state[si:si + 2], None if gi is None else m.variables[gi],
                                  None if bi is None else m.variables[bi], training=m.training)


def _dropout(m: OracleModel, rate):
    li = m.n_dropout
    m.n_dropout += 1
    return lambda x: T.dropout(x, rate, m.training, m.drop_seed, m.drop_counter, m.call_id, li)


# ---------------------------------------------------------------------------
# unet.py
# ---------------------------------------------------------------------------
def _double_conv(m, cin, f, k, norm_type, apply_dropout):
    """unet.py:20-36 (a normalization string that is neither 'batchnorm' nor 'instancenorm' adds no norm layer)."""
    ops = []
    c = cin
    for _ in range(2):
        ops.append(_conv(m, c, f, k, 1, "same", use_bias=False))
        if norm_type.lower() == 'batchnorm':
            ops.append(_bnorm(m, f, affine=True))
        elif norm_type.lower() == 'instancenorm':
            ops.append(_inorm(m, f, affine=True))
        ops.append(torch.relu)
        if apply_dropout:
            ops.append(_dropout(m, 0.5))
        c = f

    def run(x):
        for op in ops:
            x = op(x)
        return x
    return run


def strided_unet(config: Dict, dtype=torch.float32) -> OracleModel:
    """unet.py:39-78."""
    filters = config['filters']
    kernel_sizes = config['kernels']
    norm_type = config['normalization']
    output_channels = config['output_channels']
    final_activation = config['final_activation']
    norm = (lambda c: _inorm(m, c, True)) if norm_type == 'instancenorm' else (lambda c: _bnorm(m, c, True))   # unet.py:55-58

    m = OracleModel(dtype)
    up_filters = filters[::-1][:-1]
    down, c = [], 3
    for f, k in list(zip(filters, kernel_sizes))[:-1]:
        down.append((_conv(m, c, f, k, 2, "same"), norm(f)))
        c = f
    skip_ch = [f for f in filters[:-1]][::-1]
    bottom = _conv(m, c, filters[-1], kernel_sizes[-1], 2, "same")
    c = filters[-1]
    ups = []
    for f, sc, k in zip(up_filters, skip_ch, kernel_sizes[:0:-1]):
        ct = _convT(m, c, f, k, 2)
        nrm = norm(sc + f)
        ups.append((ct, nrm))
        c = sc + f
    last = _convT(m, c, output_channels, 4, 2)

    def forward(x):
        skips = []
        for cv, nrm in down:
            x = torch.relu(nrm(cv(x)))
            skips.insert(0, x)
        x = bottom(x)
        for (ct, nrm), skip in zip(ups, skips):
            x = ct(x)
            x = torch.cat([skip, x], dim=-1)
            x = torch.relu(nrm(x))
        return T.activation(last(x), final_activation)
    m.forward = forward
    return m


def unet_generator(config: Dict, dtype=torch.float32) -> OracleModel:
    """unet.py:81-124 (expansion == 'upsample' branch; the other branch cannot build, unet.py:117)."""
    filters = config['filters']
    kernel_sizes = config['kernels']
    expansion = config['expansion']
    norm_type = config['normalization']
    apply_dropout = config['dropout']
    output_channels = config['output_channels']
    final_activation = config['final_activation']
    if expansion != 'upsample':
        raise TypeError("reference unet.py:117 calls ReLU(x): the non-'upsample' branch fails to build")

    m = OracleModel(dtype)
    up_filters = filters[::-1][:-1]
    downs, c = [], 3
    for f, k in list(zip(filters, kernel_sizes))[:-1]:
        downs.append(_double_conv(m, c, f, k, norm_type, apply_dropout))
        c = f
    bottom = _double_conv(m, c, filters[-1], kernel_sizes[-1], norm_type, apply_dropout)
    c = filters[-1]
    skip_ch = filters[:-1][::-1]
    ups = []
    for f, sc, k in zip(up_filters, skip_ch, kernel_sizes[:0:-1]):
        ups.append(_double_conv(m, sc + c, f, k, norm_type, apply_dropout))
        c = f
    head = _conv(m, c, output_channels, 1, 1, "same", use_bias=True, init="glorot")

    def forward(x):
        skips = []
        for dc in downs:
            x = dc(x)
            skips.insert(0, x)
            x = T.avg_pool2(x)
        x = bottom(x)
        for dc, skip in zip(ups, skips):
            x = T.upsample2(x)
            x = torch.cat([skip, x], dim=-1)
            x = dc(x)
        return T.activation(head(x), final_activation)
    m.forward = forward
    return m


# ---------------------------------------------------------------------------
# resnet.py
# ---------------------------------------------------------------------------
def resnet_generator(config: Dict, dtype=torch.float32) -> OracleModel:
    """resnet.py:63-85 (conv7s1 :38-46, downsample :49-53, residual :26-35, upsample :56-60)."""
    f = config['filters']
    m = OracleModel(dtype)
    c7a = _conv(m, 3, f, 7, 1, "valid")
    d1 = _conv(m, f, 2 * f, 3, 2, "same")
    d2 = _conv(m, 2 * f, 4 * f, 3, 2, "same")
    res = [(_conv(m, 4 * f, 4 * f, 3, 1, "valid"), _conv(m, 4 * f, 4 * f, 3, 1, "valid")) for _ in range(9)]
    u1 = _convT(m, 4 * f, 2 * f, 3, 2)
    u2 = _convT(m, 2 * f, f, 3, 2)
    c7b = _conv(m, f, 3, 7, 1, "valid")

    def forward(x):
        x = torch.relu(T.instance_norm(c7a(T.reflection_pad(x, 3, 3))))
        x = torch.relu(T.instance_norm(d1(x)))
        x = torch.relu(T.instance_norm(d2(x)))
        for ca, cb in res:
            y = torch.relu(T.instance_norm(ca(T.reflection_pad(x, 1, 1))))
            y = T.instance_norm(cb(T.reflection_pad(y, 1, 1)))
            x = x + y
        x = torch.relu(T.instance_norm(u1(x)))
        x = torch.relu(T.instance_norm(u2(x)))
        return torch.tanh(c7b(T.reflection_pad(x, 3, 3)))
    m.forward = forward
    return m


def simple_discriminator(config: Dict, dtype=torch.float32) -> OracleModel:
    """resnet.py:87-105."""
    down_filters = config['filters']
    kernel_size = config['kernels']
    norm_type = config['normalization']
    m = OracleModel(dtype)
    convs, c = [], 3
    for k, f in zip(kernel_size, down_filters):
        cv = _conv(m, c, f, k, 2, "same")
        nrm = _inorm(m, f, False) if norm_type == 'instancenorm' else _bnorm(m, f, False)      # resnet.py:97-100
        convs.append((cv, nrm))
        c = f
    head = _conv(m, c, 1, 1, 1, "same")

    def forward(x):
        for cv, nrm in convs:
            x = T.leaky_relu(nrm(cv(x)), 0.2)
        return head(x)
    m.forward = forward
    return m


def create_model(config: Dict, dtype=torch.float32) -> OracleModel:
    """model.py:22-32: dispatch on config['type'] by builder __name__."""
    fns = [simple_discriminator, resnet_generator, unet_generator, strided_unet]
    return {fn.__name__: fn for fn in fns}[config["type"]](config, dtype)
