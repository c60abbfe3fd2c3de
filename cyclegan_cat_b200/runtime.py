"""Host-side objects that stand in for Keras models/tensors/variables.

`Model` is what the four builders return (reference: keras `Model`, unet.py:78,123,
resnet.py:85,105).  It keeps float32 master weights in Keras
`trainable_variables` order and runs forward passes through the C-ABI
(`cg_net_forward`).  torch is used only to own device memory and streams.
"""
import ctypes
from typing import List, Sequence

import numpy as np

from . import _lib, ir


def _torch():
    import torch
    return torch


def _require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.NativeError("cyclegan_cat_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    _lib.init_device(torch.cuda.current_device())
    return torch


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream_ptr(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceTensor:
    """Result of a model call: `.shape`, `.numpy()`, `[i]`, numpy arithmetic (predict.py:26-27,32-36)."""

    def __init__(self, t):
        self._t = t

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def torch(self):
        return self._t

    def numpy(self):
        return self._t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, idx):
        return self.numpy()[idx]

    def __len__(self):
        return self._t.shape[0]


class Variable:
    """One entry of `model.trainable_variables` (a view into the flat parameter buffer) or, with
    `state=True`, of `model.non_trainable_variables` (BatchNormalization moving statistics)."""

    def __init__(self, model, index, shape, offset, role, state=False):
        self._model, self.index, self.shape, self.offset, self.role = model, index, tuple(shape), offset, role
        self.size = int(np.prod(shape))
        self.state = state

    def numpy(self):
        flat = self._model._state_flat_host() if self.state else self._model._flat_host()
        return flat[self.offset:self.offset + self.size].reshape(self.shape).copy()

    def assign(self, value):
        self._model._assign(self, np.asarray(value, np.float32))

    def __repr__(self):
        return f"<Variable {self._model.name}[{self.index}] shape={self.shape}>"


def to_device_f32(x, torch):
    """Accept numpy (float64/float32/ints), torch tensors or DeviceTensor; return float32 NHWC cuda tensor."""
    if isinstance(x, DeviceTensor):
        x = x.torch
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    if x.dtype != torch.float32:
        x = x.to(torch.float32)
    if not x.is_cuda:
        x = x.pin_memory().cuda(non_blocking=True) if x.numel() > 0 else x.cuda()
    return x.contiguous()


class Model:
    def __init__(self, graph: ir.Graph, name: str, mode: str = "bf16", seed=None):
        self.graph, self.name = graph, name
        if mode not in ir.MODE_BY_NAME:
            raise ValueError(f"mode must be one of {sorted(ir.MODE_BY_NAME)}")
        self.mode = mode
        specs = graph.var_specs()
        self._specs = specs
        offs, off = [], 0
        for shape, _ in specs:
            offs.append(off)
            off += int(np.prod(shape))
        self.n_params = off
        roles = []
        for L in graph.layers:
            if L.op in (ir.OP_CONV, ir.OP_CONVT):
                roles += [0] + ([1] if L.has_bias else [])
            elif L.op in (ir.OP_INORM, ir.OP_BNORM) and L.affine:
                roles += [2, 3]
        self.trainable_variables: List[Variable] = [
            Variable(self, i, s, o, r) for i, ((s, _), o, r) in enumerate(zip(specs, offs, roles))]
        self._host = np.zeros(self.n_params, np.float32)
        self._dev = None            # flat float32 cuda tensor (master weights)
        # non-trainable state: BatchNormalization [moving_mean (0), moving_variance (1)] per layer, Keras order
        self.non_trainable_variables: List[Variable] = []
        soff = 0
        for i, (shape, _) in enumerate(graph.state_specs()):
            self.non_trainable_variables.append(Variable(self, i, shape, soff, 4 + (i & 1), state=True))
            soff += int(np.prod(shape))
        self.n_state = soff
        self._state_host = np.concatenate(
            [(np.zeros if kind == "zeros" else np.ones)(shape, np.float32).ravel()
             for shape, kind in graph.state_specs()]) if soff else np.zeros(0, np.float32)
        self._state_dev = None
        self._drop_seed = 0
        self._handle = None
        self._ws = {}
        self.owner = None           # set by CycleGan when a native trainer shares the buffer
        self.initialize(seed)

    # -- weights -------------------------------------------------------------------
    def initialize(self, seed=None):
        """Keras initializers used by the reference: random_normal(0, .02) kernels (unet.py:23,46,90;
        resnet.py:66,94), zero biases, gamma=1/beta=0, glorot-uniform 1x1 head (unet.py:121)."""
        rng = np.random.RandomState(seed)
        for v, (shape, kind) in zip(self.trainable_variables, self._specs):
            if kind == "normal":
                a = rng.normal(0.0, 0.02, size=shape)
            elif kind == "zeros":
                a = np.zeros(shape)
            elif kind == "ones":
                a = np.ones(shape)
            elif kind == "glorot":
                kh, kw, cin, cout = shape
                lim = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
                a = rng.uniform(-lim, lim, size=shape)
            else:
                raise ValueError(kind)
            self._host[v.offset:v.offset + v.size] = a.astype(np.float32).ravel()
        self._push()

    def _push(self):
        if self._dev is not None and self.n_params:
            torch = _torch()
            self._dev.copy_(torch.from_numpy(self._host))

    def _flat_host(self):
        if self._dev is not None:
            self._host = self._dev.detach().cpu().numpy()[:self.n_params]
        return self._host

    def _state_flat_host(self):
        if self._state_dev is not None:
            self._state_host = self._state_dev.detach().cpu().numpy()
        return self._state_host

    def _assign(self, var: Variable, value):
        assert value.shape == var.shape, (value.shape, var.shape)
        flat, dev = (self._state_flat_host(), self._state_dev) if var.state else (self._flat_host(), self._dev)
        flat[var.offset:var.offset + var.size] = value.ravel()
        if dev is not None:
            torch = _torch()
            dev[var.offset:var.offset + var.size].copy_(torch.from_numpy(value.ravel().copy()))

    @property
    def weights(self):
        """keras `Model.weights`: the trainable variables, then the non-trainable ones."""
        return self.trainable_variables + self.non_trainable_variables

    def get_weights(self):
        return [v.numpy() for v in self.weights]

    def set_weights(self, arrays: Sequence[np.ndarray]):
        """keras `Model.set_weights`; a list holding only the trainable variables is accepted too."""
        nt = len(self.trainable_variables)
        assert len(arrays) in (nt, nt + len(self.non_trainable_variables)), len(arrays)
        flat = self._flat_host()
        for v, a in zip(self.trainable_variables, arrays[:nt]):
            a = np.asarray(a, np.float32)
            assert a.shape == v.shape, (a.shape, v.shape)
            flat[v.offset:v.offset + v.size] = a.ravel()
        self._host = flat
        self._push()
        if len(arrays) > nt:
            st = self._state_flat_host()
            for v, a in zip(self.non_trainable_variables, arrays[nt:]):
                a = np.asarray(a, np.float32)
                assert a.shape == v.shape, (a.shape, v.shape)
                st[v.offset:v.offset + v.size] = a.ravel()
            self._state_host = st
            if self._state_dev is not None:
                self._state_dev.copy_(_torch().from_numpy(st))

    def set_dropout_seed(self, seed: int):
        """Seed of the counter-based dropout masks (also restarts the call counter)."""
        self._drop_seed = int(seed) & (2 ** 64 - 1)
        if self._handle is not None:
            _lib.check(_lib.load().cg_net_set_seed(self._handle, ctypes.c_uint64(self._drop_seed)), "cg_net_set_seed")

    # -- native handle ---------------------------------------------------------------
    def handle(self):
        """cg_net_create is pure host planning, so it also works (and validates the graph) without a GPU."""
        if self._handle is None:
            lib = _lib.load()
            h = ctypes.c_void_p()
            arr = self.graph.to_c_array()
            _lib.check(lib.cg_net_create(arr, len(self.graph.layers), ir.MODE_BY_NAME[self.mode], ctypes.byref(h)),
                       "cg_net_create")
            n = ctypes.c_int64()
            _lib.check(lib.cg_net_param_floats(h, ctypes.byref(n)), "cg_net_param_floats")
            if n.value != self.n_params:
                raise _lib.NativeError(f"parameter count mismatch host {self.n_params} vs native {n.value}")
            _lib.check(lib.cg_net_state_floats(h, ctypes.byref(n)), "cg_net_state_floats")
            if n.value != self.n_state:
                raise _lib.NativeError(f"state size mismatch host {self.n_state} vs native {n.value}")
            _lib.check(lib.cg_net_set_seed(h, ctypes.c_uint64(self._drop_seed)), "cg_net_set_seed")
            self._handle = h
        return self._handle

    def device_params(self):
        torch = _require_cuda()
        if self._dev is None:      # a graph without variables still hands the C-ABI a valid (unused) pointer
            self._dev = torch.from_numpy(self._host).cuda() if self.n_params else torch.zeros(4, device="cuda")
        self.device_state()
        return self._dev

    def device_state(self):
        """The moving statistics live in a caller-owned device buffer bound to the handle (cg_net_bind_state)."""
        if self.n_state and self._state_dev is None:
            torch = _require_cuda()
            self._state_dev = torch.from_numpy(self._state_host).cuda()
            _lib.check(_lib.load().cg_net_bind_state(self.handle(), _ptr(self._state_dev)), "cg_net_bind_state")
        return self._state_dev

    def out_shape(self, N, H, W):
        out = (ctypes.c_int * 4)()
        _lib.check(_lib.load().cg_net_out_shape(self.handle(), N, H, W, ctypes.byref(out)), "cg_net_out_shape")
        return tuple(out)

    # -- Keras Model surface -----------------------------------------------------------
    def __call__(self, x, training=False):
        torch = _require_cuda()
        lib = _lib.load()
        xd = to_device_f32(x, torch)
        cin = self.graph.channels[0]
        if xd.dim() != 4 or xd.shape[3] != cin:
            raise ValueError(f"expected NHWC input with {cin} channels, got {tuple(xd.shape)}")
        N, H, W, _ = xd.shape
        f = self.graph.down_factor()
        if H % f or W % f:
            raise ValueError(f"H, W must be multiples of {f} for this model, got {(H, W)}")
        key = (N, H, W)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = ctypes.c_size_t()
            _lib.check(lib.cg_net_workspace_bytes(self.handle(), N, H, W, 0, ctypes.byref(nbytes)),
                       "cg_net_workspace_bytes")
            self._ws.clear()
            ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device="cuda")
            self._ws[key] = ws
        y = torch.empty(self.out_shape(N, H, W), dtype=torch.float32, device="cuda")
        _lib.check(lib.cg_net_set_training(self.handle(), int(bool(training))), "cg_net_set_training")
        _lib.check(lib.cg_net_forward(self.handle(), _ptr(self.device_params()), _ptr(xd), _ptr(y), _ptr(ws),
                                      ws.numel(), N, H, W, 0, _stream_ptr(torch)), "cg_net_forward")
        return DeviceTensor(y)

    def intermediates(self):
        """{tensor id: float32 NHWC numpy array} of every tensor the last `model(x)` / cg_net_forward materialised
        (tensor 0 = input, i + 1 = output of layer i; fused layers have no tensor of their own).  The Keras
        counterpart is `keras.Model(model.input, [l.output for l in model.layers])`; the parity tests feed these to the
        layer-by-layer oracle (cg_net_fetch_tensor)."""
        return _fetch_all(lambda t, out, shape, st: _lib.load().cg_net_fetch_tensor(self.handle(), t, out, shape, st),
                          len(self.graph.layers) + 1)

    def predict(self, x, batch_size=32):
        """keras Model.predict (model.py:268-269): batched forward, numpy result."""
        x = np.asarray(x) if not _torch().is_tensor(x) and not isinstance(x, DeviceTensor) else x
        outs = []
        n = len(x)
        for i in range(0, n, batch_size):
            outs.append(self(x[i:i + batch_size]).numpy())
        return np.concatenate(outs, 0)

    def __del__(self):
        try:
            if self._handle is not None and self.owner is None:
                _lib.load().cg_net_destroy(self._handle)
        except Exception:
            pass


def _fetch_all(fetch, n_tensors):
    torch = _require_cuda()
    st = _stream_ptr(torch)
    out = {}
    for t in range(n_tensors):
        shape = (ctypes.c_int * 4)()
        rc = fetch(t, ctypes.c_void_p(0), ctypes.byref(shape), st)
        if rc == 1:
            continue            # fused away: never materialised
        _lib.check(rc, "fetch_tensor (query)")
        buf = torch.empty(tuple(shape), dtype=torch.float32, device="cuda")
        _lib.check(fetch(t, _ptr(buf), ctypes.byref(shape), st), "fetch_tensor")
        out[t] = buf.cpu().numpy()
    return out


def reflection_pad_device(x, pad: int):
    """ReflectionPadding2D()(x) (resnet.py:21-23) on the GPU; integer inputs stay integer
    (exactly representable in fp32 for |v| < 2^24), test vector unittests/test_resnet.py:31-47."""
    torch = _require_cuda()
    a = np.asarray(x.numpy() if isinstance(x, DeviceTensor) else x)
    if a.ndim != 4:
        raise ValueError("ReflectionPadding2D expects a 4-D NHWC input")
    N, H, W, C = a.shape
    if pad >= H or pad >= W:
        raise ValueError("reflect padding must be smaller than the input")
    g = ir.Graph(channels=[C])
    g.reflect_pad(g.input, pad)
    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.cg_net_create(g.to_c_array(), 1, ir.MODE_FP32_CHECK, ctypes.byref(h)), "cg_net_create")
    try:
        xd = torch.from_numpy(np.ascontiguousarray(a.astype(np.float32))).cuda()
        nbytes = ctypes.c_size_t()
        _lib.check(lib.cg_net_workspace_bytes(h, N, H, W, 0, ctypes.byref(nbytes)), "cg_net_workspace_bytes")
        ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device="cuda")
        y = torch.empty((N, H + 2 * pad, W + 2 * pad, C), dtype=torch.float32, device="cuda")
        dummy = torch.zeros(4, dtype=torch.float32, device="cuda")
        _lib.check(lib.cg_net_forward(h, _ptr(dummy), _ptr(xd), _ptr(y), _ptr(ws), ws.numel(), N, H, W, 0,
                                      _stream_ptr(torch)), "cg_net_forward")
        out = y.cpu().numpy()
    finally:
        lib.cg_net_destroy(h)

    class _Padded(DeviceTensor):
        def numpy(self_inner):
            return out.astype(a.dtype) if np.issubdtype(a.dtype, np.integer) else out
    return _Padded(y)
