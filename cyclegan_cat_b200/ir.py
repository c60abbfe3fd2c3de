"""Layer-graph IR shared by the four model builders and the C-ABI.

A `Graph` is the forward program of one Keras functional model of the
reference (`Model(inputs, outputs)`, unet.py:78,123 / resnet.py:85,105) written
as a list of `cg_layer_desc` nodes (include/cyclegan_b200.h).  Tensor 0 is the
input; layer i produces tensor i+1.  Variables are enumerated in Keras
`trainable_variables` order.
"""
import ctypes
from dataclasses import dataclass, field
from typing import List

# cg_op
(OP_CONV, OP_CONVT, OP_INORM, OP_ACT, OP_RPAD, OP_ADD, OP_CONCAT, OP_AVGPOOL, OP_UPSAMPLE, OP_BNORM,
 OP_DROPOUT) = range(1, 12)
# cg_act
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH, ACT_SIGMOID = range(5)
ACT_BY_NAME = {None: ACT_NONE, "linear": ACT_NONE, "relu": ACT_RELU, "tanh": ACT_TANH, "sigmoid": ACT_SIGMOID}
# cg_loss
LOSS_BY_NAME = {"mse": 0, "mae": 1, "bce": 2}
# cg_mode
MODE_BF16, MODE_FP32_CHECK = 0, 1
MODE_BY_NAME = {"bf16": MODE_BF16, "fp32": MODE_FP32_CHECK}
# cg_opt
OPT_ADAM, OPT_SGD, OPT_RMSPROP, OPT_ADABELIEF = range(4)


class LayerDesc(ctypes.Structure):
    """ctypes mirror of `cg_layer_desc`."""
    _fields_ = [("op", ctypes.c_int32), ("in0", ctypes.c_int32), ("in1", ctypes.c_int32),
                ("cin", ctypes.c_int32), ("cout", ctypes.c_int32), ("k", ctypes.c_int32),
                ("stride", ctypes.c_int32), ("same", ctypes.c_int32), ("has_bias", ctypes.c_int32),
                ("act", ctypes.c_int32), ("affine", ctypes.c_int32), ("pad", ctypes.c_int32),
                ("eps", ctypes.c_float), ("slope", ctypes.c_float), ("momentum", ctypes.c_float),
                ("rate", ctypes.c_float)]


class VarInfo(ctypes.Structure):
    """ctypes mirror of `cg_var_info`."""
    _fields_ = [("layer", ctypes.c_int32), ("role", ctypes.c_int32), ("ndim", ctypes.c_int32),
                ("shape", ctypes.c_int32 * 4), ("offset", ctypes.c_int64)]


class AdamCfg(ctypes.Structure):
    _fields_ = [("learning_rate", ctypes.c_float), ("beta_1", ctypes.c_float),
                ("beta_2", ctypes.c_float), ("epsilon", ctypes.c_float), ("kind", ctypes.c_int32)]


class TrainCfg(ctypes.Structure):
    _fields_ = [("loss", ctypes.c_int32), ("w_cycle", ctypes.c_float), ("w_identity", ctypes.c_float),
                ("w_generator", ctypes.c_float), ("w_discriminator", ctypes.c_float),
                ("adam", AdamCfg * 4)]


@dataclass
class Layer:
    op: int
    in0: int
    in1: int = -1
    cin: int = 0
    cout: int = 0
    k: int = 0
    stride: int = 1
    same: int = 0
    has_bias: int = 0
    act: int = 0
    affine: int = 0
    pad: int = 0
    eps: float = 1e-3
    slope: float = 0.2
    momentum: float = 0.99
    rate: float = 0.0
    init: str = "normal"          # host-side only: kernel initializer kind

    def to_c(self) -> LayerDesc:
        return LayerDesc(self.op, self.in0, self.in1, self.cin, self.cout, self.k, self.stride, self.same,
                         self.has_bias, self.act, self.affine, self.pad, self.eps, self.slope, self.momentum,
                         self.rate)


@dataclass
class Graph:
    """Forward program + channel bookkeeping; the builder methods read like Keras layer calls."""
    layers: List[Layer] = field(default_factory=list)
    channels: List[int] = field(default_factory=lambda: [3])      # per tensor id

    def _emit(self, layer: Layer, cout: int) -> int:
        self.layers.append(layer)
        self.channels.append(cout)
        return len(self.layers)

    @property
    def input(self) -> int:
        return 0

    @property
    def output(self) -> int:
        return len(self.layers)

    # -- Keras-like layer calls ------------------------------------------------
    def conv(self, x, filters, k, stride=1, padding="same", use_bias=True, init="normal"):
        cin = self.channels[x]
        return self._emit(Layer(OP_CONV, x, cin=cin, cout=filters, k=k, stride=stride,
                                same=int(padding == "same"), has_bias=int(use_bias), init=init), filters)

    def conv_transpose(self, x, filters, k, stride=2):
        cin = self.channels[x]
        return self._emit(Layer(OP_CONVT, x, cin=cin, cout=filters, k=k, stride=stride, same=1, has_bias=1), filters)

    def instance_norm(self, x, affine=True, eps=1e-3):
        c = self.channels[x]
        return self._emit(Layer(OP_INORM, x, cin=c, cout=c, affine=int(affine), eps=eps), c)

    def batch_norm(self, x, affine=True, eps=1e-3, momentum=0.99):
        """keras BatchNormalization(): momentum 0.99, epsilon 1e-3; center/scale = `affine`."""
        c = self.channels[x]
        return self._emit(Layer(OP_BNORM, x, cin=c, cout=c, affine=int(affine), eps=eps, momentum=momentum), c)

    def dropout(self, x, rate=0.5):
        c = self.channels[x]
        return self._emit(Layer(OP_DROPOUT, x, cin=c, cout=c, rate=rate), c)

    def act(self, x, kind, slope=0.2):
        c = self.channels[x]
        if kind == ACT_NONE:
            return x
        return self._emit(Layer(OP_ACT, x, cin=c, cout=c, act=kind, slope=slope), c)

    def reflect_pad(self, x, pad):
        c = self.channels[x]
        return self._emit(Layer(OP_RPAD, x, cin=c, cout=c, pad=pad), c)

    def add(self, a, b):
        c = self.channels[a]
        assert c == self.channels[b]
        return self._emit(Layer(OP_ADD, a, b, cin=c, cout=c), c)

    def concat(self, a, b):
        ca, cb = self.channels[a], self.channels[b]
        return self._emit(Layer(OP_CONCAT, a, b, cin=ca, cout=ca + cb), ca + cb)

    def avg_pool(self, x):
        c = self.channels[x]
        return self._emit(Layer(OP_AVGPOOL, x, cin=c, cout=c), c)

    def upsample(self, x):
        c = self.channels[x]
        return self._emit(Layer(OP_UPSAMPLE, x, cin=c, cout=c), c)

    # -- derived ----------------------------------------------------------------
    def var_specs(self):
        """[(shape, init kind)] in Keras trainable_variables order."""
        specs = []
        for L in self.layers:
            if L.op == OP_CONV:
                specs.append(((L.k, L.k, L.cin, L.cout), L.init))
                if L.has_bias:
                    specs.append(((L.cout,), "zeros"))
            elif L.op == OP_CONVT:
                specs.append(((L.k, L.k, L.cout, L.cin), L.init))
                if L.has_bias:
                    specs.append(((L.cout,), "zeros"))
            elif L.op in (OP_INORM, OP_BNORM) and L.affine:
                specs.append(((L.cin,), "ones"))
                specs.append(((L.cin,), "zeros"))
        return specs

    def state_specs(self):
        """[(shape, init kind)] of the non-trainable variables in Keras order: per BatchNormalization
        [moving_mean (zeros), moving_variance (ones)]."""
        specs = []
        for L in self.layers:
            if L.op == OP_BNORM:
                specs.append(((L.cin,), "zeros"))
                specs.append(((L.cin,), "ones"))
        return specs

    def has_dropout(self) -> bool:
        return any(L.op == OP_DROPOUT for L in self.layers)

    def to_c_array(self):
        arr = (LayerDesc * len(self.layers))()
        for i, L in enumerate(self.layers):
            arr[i] = L.to_c()
        return arr

    def down_factor(self) -> int:
        """Input H, W must be divisible by this (number of stride-2 / pool stages)."""
        f = 1
        best = 1
        # walk the main chain conservatively: count downsamplings
        for L in self.layers:
            if (L.op == OP_CONV and L.stride == 2) or L.op == OP_AVGPOOL:
                f *= 2
                best = max(best, f)
            elif (L.op == OP_CONVT and L.stride == 2) or L.op == OP_UPSAMPLE:
                f = max(1, f // 2)
        return best

    def flops(self, H: int, W: int) -> float:
        """Algorithmic forward FLOPs per image: conv 2*Ho*Wo*Cout*k^2*Cin,
        convT 2*Hi*Wi*Cin*k^2*Cout."""
        hw = [(H, W)]
        total = 0.0
        for L in self.layers:
            h, w = hw[L.in0]
            if L.op == OP_CONV:
                if L.same:
                    ho, wo = -(-h // L.stride), -(-w // L.stride)
                else:
                    ho, wo = (h - L.k) // L.stride + 1, (w - L.k) // L.stride + 1
                total += 2.0 * ho * wo * L.cout * L.k * L.k * L.cin
                hw.append((ho, wo))
            elif L.op == OP_CONVT:
                total += 2.0 * h * w * L.cin * L.k * L.k * L.cout
                hw.append((h * L.stride, w * L.stride))
            elif L.op == OP_RPAD:
                hw.append((h + 2 * L.pad, w + 2 * L.pad))
            elif L.op == OP_AVGPOOL:
                hw.append((h // 2, w // 2))
            elif L.op == OP_UPSAMPLE:
                hw.append((h * 2, w * 2))
            else:
                hw.append((h, w))
        return total

    def conv_out_elems(self, H: int, W: int) -> int:
        """Conv / transposed-conv output elements per image (the unit of the HBM-side roofline)."""
        hw = [(H, W)]
        total = 0
        for L in self.layers:
            h, w = hw[L.in0]
            if L.op == OP_CONV:
                ho = -(-h // L.stride) if L.same else (h - L.k) // L.stride + 1
                wo = -(-w // L.stride) if L.same else (w - L.k) // L.stride + 1
                total += ho * wo * L.cout
                hw.append((ho, wo))
            elif L.op == OP_CONVT:
                total += h * L.stride * w * L.stride * L.cout
                hw.append((h * L.stride, w * L.stride))
            elif L.op == OP_RPAD:
                hw.append((h + 2 * L.pad, w + 2 * L.pad))
            elif L.op == OP_AVGPOOL:
                hw.append((h // 2, w // 2))
            elif L.op == OP_UPSAMPLE:
                hw.append((h * 2, w * 2))
            else:
                hw.append((h, w))
        return total

    def first_layer_flops(self, H: int, W: int) -> float:
        """FLOPs of the first conv, whose data-gradient is never needed."""
        hw = [(H, W)]
        for L in self.layers:
            h, w = hw[L.in0]
            if L.op == OP_CONV:
                ho = -(-h // L.stride) if L.same else (h - L.k) // L.stride + 1
                wo = -(-w // L.stride) if L.same else (w - L.k) // L.stride + 1
                return 2.0 * ho * wo * L.cout * L.k * L.k * L.cin
            if L.op == OP_RPAD:
                hw.append((h + 2 * L.pad, w + 2 * L.pad))
            else:
                hw.append((h, w))
        return 0.0
