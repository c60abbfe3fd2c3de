"""`CycleGan` trainer with the reference's interface (cyclegan/model.py:22-362).

`create_model`, `accuracy`, `CycleGan.__init__/build_models/validate_step/train_step/
train/save_model/load_model` keep the reference's names, arguments and metric keys.
`train_step` is ONE native call (`cg_train_step`): 6 generator forwards, 4 discriminator
forwards, the hand-scheduled backward of SURVEY.md 3.2, and four fused Adam updates.
"""
import ctypes
import logging
import os
from os.path import join
from typing import Dict

import numpy as np

from .. import _lib, ir
from ..model_processing.load_model import Bunch, namespace2yaml
from ..runtime import DeviceTensor, Model, _ptr, _require_cuda, _stream_ptr, to_device_f32
from .losses import get_loss_obj
from .optimizers import get_optimizer
from .resnet import resnet_generator, simple_discriminator
from .unet import strided_unet, unet_generator

logger = logging.getLogger(__name__)
logger.setLevel(logging.INFO)

METRIC_KEYS = ("gAB_loss", "gBA_loss", "dA_loss", "dB_loss", "dA_acc", "dB_acc")   # model.py:126-133
NET_NAMES = ("g_AB", "g_BA", "d_A", "d_B")


def create_model(config: Dict, mode: str = "bf16") -> Model:
    """model.py:22-32: dispatch on config['type'] by builder __name__ (KeyError if unknown)."""
    chosen_type = config["type"]
    MODEL_FUNCTION = [simple_discriminator, resnet_generator, unet_generator, strided_unet]
    model_type_map = {model.__name__: model for model in MODEL_FUNCTION}
    return model_type_map[chosen_type](config, mode=mode)


def accuracy(real, fake):
    """model.py:35-54 on host arrays (the train step computes it natively)."""
    r = np.asarray(real.numpy() if hasattr(real, "numpy") else real, np.float32)
    f = np.asarray(fake.numpy() if hasattr(fake, "numpy") else fake, np.float32)
    predictions = (np.concatenate([r, f], 0) > 0.5).astype(np.float32)
    labels = np.concatenate([np.ones_like(r), np.zeros_like(f)], 0)
    return np.float32((predictions == labels).astype(np.float32).mean())


class _MetricBuf:
    """The metric vector one step wrote to device memory, copied to the host at most once and only when someone looks."""

    def __init__(self, dev):
        self.dev, self.host = dev, None

    def get(self):
        if self.host is None:
            self.host = self.dev.cpu().numpy()
            self.dev = None
        return self.host

    @staticmethod
    def fetch_many(bufs):
        """One synchronisation and one device-to-host copy for any number of pending steps."""
        todo = [b for b in {id(b): b for b in bufs}.values() if b.host is None]
        if len(todo) == 1:
            todo[0].get()
        elif todo:
            import torch
            host = torch.stack([b.dev for b in todo]).cpu().numpy()
            for b, h in zip(todo, host):
                b.host, b.dev = h, None


class Scalar:
    """A metric living in device memory; `.numpy()` / float() sync lazily (the reference syncs every
    step at model.py:301 -- this does not, until someone looks)."""

    def __init__(self, buf, i):
        self._buf, self._i = buf if isinstance(buf, _MetricBuf) else _MetricBuf(buf), i

    def numpy(self):
        return np.float32(self._buf.get()[self._i])

    def __float__(self):
        return float(self._buf.get()[self._i])

    def __repr__(self):
        return f"Scalar({float(self):.6g})"


class Mean:
    """keras.metrics.Mean stand-in (model.py:166-183).  `update_state` with a device `Scalar` only queues it: nothing
    is read back until `result()` (the reference's `display_metrics` forces a device->host sync every step, model.py:301;
    here the steps of one progress-bar refresh are fetched together)."""

    def __init__(self, name=None):
        self.name, self.total, self.count, self._pending = name, 0.0, 0, []

    def update_state(self, v):
        if isinstance(v, Scalar):
            self._pending.append(v)
        else:
            self.total += float(v)
        self.count += 1

    def result(self):
        if self._pending:
            _MetricBuf.fetch_many([s._buf for s in self._pending])
            self.total += float(sum(float(s) for s in self._pending))
            self._pending = []
        return np.float32(self.total / max(self.count, 1))

    def reset_states(self):
        self.total, self.count, self._pending = 0.0, 0, []


class CycleGan:
    def __init__(self, model_config: Bunch, train_config: Bunch = None, mode: str = "bf16",
                 summaries: bool = True, display_interval: float = 0.5):
        self.model_config = model_config
        self.mode = mode
        self.model_folder = join(self.model_config.location, self.model_config.name)
        # model.py:62-66: the reference always creates the two TensorBoard writers; here they are created on first use
        # (torch's SummaryWriter), so a CycleGan that never calls train() leaves no event files behind
        self._summaries = summaries
        self._writers = {}
        self.display_interval = display_interval        # seconds between progress-bar metric refreshes (each one syncs)
        self._last_display = 0.0
        self.train_config = train_config
        self.g_AB_optimizer = get_optimizer(self.train_config.g_opt)      # model.py:68-71
        self.g_BA_optimizer = get_optimizer(self.train_config.g_opt)
        self.d_A_optimizer = get_optimizer(self.train_config.d_opt)
        self.d_B_optimizer = get_optimizer(self.train_config.d_opt)

        self.loss_weights = self.model_config.loss_weights
        self._trainer = None
        self._bound_shape = None
        self._world = 1
        self.build_models()
        if self.model_config.new:                                          # model.py:75-78
            self.model_config.new = False
        else:
            self.load_model()

    def _writer(self, kind):
        if not self._summaries:
            return None
        if kind not in self._writers:
            from torch.utils.tensorboard import SummaryWriter
            self._writers[kind] = SummaryWriter(join(self.model_folder, kind))
        return self._writers[kind]

    @property
    def train_summaries(self):
        return self._writer("train")

    @property
    def val_summaries(self):
        return self._writer("validation")

    def build_models(self):
        gen_config = self.model_config.generator
        disc_config = self.model_config.discriminator
        self.g_AB = create_model(gen_config, self.mode)
        self.g_BA = create_model(gen_config, self.mode)
        self.d_A = create_model(disc_config, self.mode)
        self.d_B = create_model(disc_config, self.mode)
        self.loss_obj = get_loss_obj(self.model_config.loss)

    # -- native trainer ----------------------------------------------------------------
    def _nets(self):
        return [self.g_AB, self.g_BA, self.d_A, self.d_B]

    def _opts(self):
        return [self.g_AB_optimizer, self.g_BA_optimizer, self.d_A_optimizer, self.d_B_optimizer]

    def _ensure_trainer(self, B, H, W):
        torch = _require_cuda()
        lib = _lib.load()
        if self._trainer is None:
            cfg = ir.TrainCfg()
            cfg.loss = self.loss_obj.kind
            w = self.loss_weights
            cfg.w_cycle, cfg.w_identity = w["cycle"], w["identity"]
            cfg.w_generator, cfg.w_discriminator = w["generator"], w["discriminator"]
            for i, o in enumerate(self._opts()):
                cfg.adam[i] = ir.AdamCfg(o.learning_rate, o.beta_1, o.beta_2, o.epsilon, o.kind)
                o._binding = (self, i)
            h = ctypes.c_void_p()
            nets = self._nets()
            _lib.check(lib.cg_trainer_create(nets[0].handle(), nets[1].handle(), nets[2].handle(),
                                             nets[3].handle(), ctypes.byref(cfg), ctypes.byref(h)),
                       "cg_trainer_create")
            self._trainer = h
            self._params = [n.device_params() for n in nets]
            self._grads = [torch.zeros_like(p) for p in self._params]
            self._m = [torch.zeros_like(p) for p in self._params]
            self._v = [torch.zeros_like(p) for p in self._params]
            self._metrics = torch.zeros(8, dtype=torch.float32, device="cuda")
            self._ws = None
            self._ws_cap = (0, 0, 0)
            # resume (model.py:335-338,344-362): load_model() read the four `<name>_optimizer.npy` files before any
            # device buffer existed; now that the slots do, they go in -- `CycleGan(new=False).train()` continues with
            # the saved iteration counts and moments, no extra call needed
            self.restore_optimizers()
        cap = self._ws_cap
        if self._ws is None or B > cap[0] or (H, W) != cap[1:]:
            nbytes = ctypes.c_size_t()
            _lib.check(lib.cg_trainer_workspace_bytes(self._trainer, B, H, W, ctypes.byref(nbytes)),
                       "cg_trainer_workspace_bytes")
            self._ws = None
            torch.cuda.empty_cache()
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
            self._ws_cap = (B, H, W)
            arr = lambda ts: (ctypes.c_void_p * 4)(*[t.data_ptr() for t in ts])
            _lib.check(lib.cg_trainer_bind(self._trainer, ctypes.byref(arr(self._params)),
                                           ctypes.byref(arr(self._grads)), ctypes.byref(arr(self._m)),
                                           ctypes.byref(arr(self._v)), _ptr(self._ws), self._ws.numel()),
                       "cg_trainer_bind")
        return torch, lib

    def _prep(self, real_a, real_b):
        torch = _require_cuda()
        a, b = to_device_f32(real_a, torch), to_device_f32(real_b, torch)
        if a.shape != b.shape or a.dim() != 4 or a.shape[3] != 3:
            raise ValueError(f"expected two NHWC batches of equal shape with 3 channels, got {tuple(a.shape)} "
                             f"and {tuple(b.shape)}")
        return a, b

    def _run(self, fn_name, real_a, real_b):
        a, b = self._prep(real_a, real_b)
        B, H, W, _ = a.shape
        torch, lib = self._ensure_trainer(B, H, W)
        out = torch.empty(8, dtype=torch.float32, device="cuda")
        _lib.check(getattr(lib, fn_name)(self._trainer, _ptr(a), _ptr(b), B, H, W, _ptr(out),
                                         _stream_ptr(torch)), fn_name)
        self._last_inputs = (a, b)      # keep alive until the stream has consumed them
        buf = _MetricBuf(out)
        return {k: Scalar(buf, i) for i, k in enumerate(METRIC_KEYS)}

    # -- reference surface ---------------------------------------------------------------
    def validate_step(self, real_a, real_b, training: bool = False) -> Dict:
        """model.py:91-134 as the reference calls it outside the tape (model.py:221: training=False): moving
        statistics for BatchNormalization, no dropout.  Instance norm is identical for training=True/False; with
        BatchNormalization / Dropout nets, training=True (a forward that also moves the moving averages but
        takes no gradient) is what `compute_gradients` runs, so it is served from there."""
        if training and any(n.n_state or n.graph.has_dropout() for n in self._nets()):
            return self._run("cg_trainer_compute_gradients", real_a, real_b)
        return self._run("cg_validate_step", real_a, real_b)

    def train_step(self, real_a, real_b) -> Dict:
        """model.py:136-154."""
        return self._run("cg_train_step", real_a, real_b)

    def compute_gradients(self, real_a, real_b):
        """The tape half of train_step (model.py:138-147): metrics + per-variable gradients, no update."""
        m = self._run("cg_trainer_compute_gradients", real_a, real_b)
        grads = {}
        for name, net, g in zip(NET_NAMES, self._nets(), self._grads):
            flat = g.detach().cpu().numpy()
            grads[name] = [flat[v.offset:v.offset + v.size].reshape(v.shape).copy() for v in net.trainable_variables]
        return m, grads

    def apply_gradients(self):
        torch = _require_cuda()
        _lib.check(_lib.load().cg_trainer_apply_gradients(self._trainer, _stream_ptr(torch)),
                   "cg_trainer_apply_gradients")

    def fetch_image(self, which: str):
        names = ("fake_b", "same_b", "fake_a", "same_a", "cycled_a", "cycled_b")
        torch = _require_cuda()
        B, H, W = self._last_inputs[0].shape[:3]
        out = torch.empty((B, H, W, 3), dtype=torch.float32, device="cuda")
        _lib.check(_lib.load().cg_trainer_fetch_image(self._trainer, names.index(which), _ptr(out),
                                                      _stream_ptr(torch)), "cg_trainer_fetch_image")
        return DeviceTensor(out)

    CALLS = ("g_AB([a;b])", "g_BA([b;a])", "g_BA(fake_b)", "g_AB(fake_a)", "d_A([a;fake_a])", "d_B([b;fake_b])")

    def call_intermediates(self, call: int):
        """{tensor id: float32 NHWC array} of model call `call` (index into CALLS) of the last step: the probe the
        layer-by-layer parity tests use (cg_trainer_fetch_tensor)."""
        from ..runtime import _fetch_all
        net = (self.g_AB, self.g_BA, self.g_BA, self.g_AB, self.d_A, self.d_B)[call]
        lib = _lib.load()
        return _fetch_all(lambda t, out, shape, st: lib.cg_trainer_fetch_tensor(self._trainer, call, t, out, shape, st),
                          len(net.graph.layers) + 1)

    def enable_data_parallel(self):
        """New functionality (the reference is single-device, train.py:36-43): one process per GPU,
        gradients all-reduced with NCCL inside the native step.  Needs torch.distributed initialised."""
        import torch.distributed as dist
        torch = _require_cuda()
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        rank, world = dist.get_rank(), dist.get_world_size()
        # identical weights everywhere: broadcast rank 0's parameters
        for n in self._nets():
            dist.broadcast(n.device_params(), src=0)
        if self._trainer is None:
            raise RuntimeError("call enable_data_parallel() after the first _ensure_trainer/step shape is known; "
                               "use prepare(B, H, W) first")
        from ..parallel import exchange_unique_id

        def make_id():
            buf = (ctypes.c_char * 128)()
            _lib.check(_lib.load().cg_comm_unique_id(buf), "cg_comm_unique_id")
            return bytes(buf)
        raw = exchange_unique_id(make_id, dist, device="cuda")
        buf = (ctypes.c_char * 128)(*raw)
        # every rank trains on its own shard: give it its own dropout stream too (before comm_init, which drops the
        # captured graphs -- the seed is baked into them)
        for n in self._nets():
            if n.graph.has_dropout():
                n.set_dropout_seed(n._drop_seed + 0x9E3779B9 * rank)
        _lib.check(_lib.load().cg_trainer_comm_init(self._trainer, buf, rank, world), "cg_trainer_comm_init")
        self._world = world

    def prepare(self, B, H, W):
        self._ensure_trainer(B, H, W)

    # optimizer state access used by optimizers.Adam ------------------------------------------
    def _get_iterations(self, i):
        it = (ctypes.c_int64 * 4)()
        _lib.check(_lib.load().cg_trainer_get_iterations(self._trainer, ctypes.byref(it)), "get_iterations")
        return int(it[i])

    def _set_iterations(self, i, value):
        it = (ctypes.c_int64 * 4)()
        _lib.check(_lib.load().cg_trainer_get_iterations(self._trainer, ctypes.byref(it)), "get_iterations")
        it[i] = int(value)
        _lib.check(_lib.load().cg_trainer_set_iterations(self._trainer, ctypes.byref(it)), "set_iterations")

    def _slot_buffers(self, i):
        """Device buffers behind the Keras slots of optimizer i, in slot creation order: Adam / AdaBelief (m, v);
        RMSprop keeps `rms` in the v buffer; SGD has none."""
        slots = self._opts()[i].slots
        return [self._v[i]] if slots == ("rms",) else [self._m[i], self._v[i]][:len(slots)]

    def _optimizer_get_weights(self, i):
        net = self._nets()[i]
        sl = lambda f: [f[x.offset:x.offset + x.size].reshape(x.shape).copy() for x in net.trainable_variables]
        out = [np.int64(self._get_iterations(i))]
        for buf in self._slot_buffers(i):
            out += sl(buf.cpu().numpy())
        return out

    def _optimizer_set_weights(self, i, weights):
        torch = _require_cuda()
        net = self._nets()[i]
        n = len(net.trainable_variables)
        bufs = self._slot_buffers(i)
        assert len(weights) == 1 + len(bufs) * n, (len(weights), len(bufs), n)
        it = (ctypes.c_int64 * 4)()
        _lib.check(_lib.load().cg_trainer_get_iterations(self._trainer, ctypes.byref(it)), "get_iterations")
        it[i] = int(weights[0])
        _lib.check(_lib.load().cg_trainer_set_iterations(self._trainer, ctypes.byref(it)), "set_iterations")
        for j, dst in enumerate(bufs):
            part = weights[1 + j * n:1 + (j + 1) * n]
            flat = np.concatenate([np.asarray(a, np.float32).ravel() for a in part]) if n else np.zeros(0, np.float32)
            dst.copy_(torch.from_numpy(flat))

    # -- epoch loop, summaries, checkpoints (host-side only) ---------------
    @staticmethod
    def _batches(dataset, batch_size):
        """`dataset.batch(batch_size)` without drop_remainder (model.py:197-198)."""
        buf_a, buf_b = [], []
        for a, b in dataset:
            buf_a.append(np.asarray(a, np.float32))
            buf_b.append(np.asarray(b, np.float32))
            if len(buf_a) == batch_size:
                yield np.stack(buf_a), np.stack(buf_b)
                buf_a, buf_b = [], []
        if buf_a:
            yield np.stack(buf_a), np.stack(buf_b)

    def train(self, train_dataset, validation_dataset):
        """model.py:156-231 with python iterables of (a, b) HWC float32 samples."""
        import tqdm
        batch_size = self.train_config.batch_size
        epochs = self.train_config.epochs
        save_images_every = self.train_config.summary["images"]
        tensorboard_samples = self.train_config.summary["samples"]
        save_model_every = self.train_config.summary["model"]
        metric_names = ["dA_loss", "dB_loss", "gAB_loss", "gBA_loss", "dA_acc", "dB_acc"]
        train_metrics_dict = {m: Mean(name=m) for m in metric_names}
        validation_metrics_dict = {m: Mean(name=m) for m in metric_names}
        train_dataset, validation_dataset = list(train_dataset), list(validation_dataset)

        if not hasattr(self, "a_samples") and not hasattr(self, "b_samples"):
            samples = validation_dataset[:tensorboard_samples]
            self.a_samples = np.stack([np.asarray(s[0], np.float32) for s in samples])
            self.b_samples = np.stack([np.asarray(s[1], np.float32) for s in samples])
            if self.val_summaries is not None:
                self.val_summaries.add_images("A", (self.a_samples + 1) / 2, 0, dataformats="NHWC")
                self.val_summaries.add_images("B", (self.b_samples + 1) / 2, 0, dataformats="NHWC")

        current_epoch = 0
        if hasattr(self.model_config, "current_epoch"):
            current_epoch = self.model_config.current_epoch

        for e in range(current_epoch, current_epoch + epochs):
            train_bar = tqdm.tqdm(self._batches(train_dataset, batch_size), desc=f"Epoch {e + 1} training", ncols=0,
                                  total=-(-len(train_dataset) // batch_size))
            for (images_a, images_b) in train_bar:
                losses = self.train_step(images_a, images_b)
                self.update_metrics(train_metrics_dict, losses)
                self.display_metrics(train_metrics_dict, train_bar)
            self.display_metrics(train_metrics_dict, train_bar, force=True)
            self.write_summaries(self.train_summaries, e, train_metrics_dict)
            if e % save_images_every == 0:
                self.write_images(e, self.a_samples, self.b_samples, tensorboard_samples)

            val_bar = tqdm.tqdm(self._batches(validation_dataset, batch_size), desc=f"Epoch {e + 1} validation",
                                ncols=0, total=-(-len(validation_dataset) // batch_size))
            for (images_a, images_b) in val_bar:
                losses = self.validate_step(images_a, images_b, training=False)
                self.update_metrics(validation_metrics_dict, losses)
                self.display_metrics(validation_metrics_dict, val_bar)
            self.display_metrics(validation_metrics_dict, val_bar, force=True)
            self.write_summaries(self.val_summaries, e, validation_metrics_dict)
            if e % save_model_every == 0:
                self.save_model()

        self.model_config.current_epoch = current_epoch + epochs
        os.makedirs(self.model_folder, exist_ok=True)
        namespace2yaml(join(self.model_folder, "model_config.yaml"), self.model_config)
        self.save_model()

    def write_summaries(self, summaries, epoch: int, metrics_dict):
        for name, metric in metrics_dict.items():
            if summaries is not None:
                summaries.add_scalar(name, float(metric.result()), epoch)
            metrics_dict[name].reset_states()

    def write_images(self, epoch: int, a_samples, b_samples, num_samples: int):
        if self.val_summaries is None:
            return
        prediction_ab = self.g_AB.predict(x=a_samples, batch_size=1)
        prediction_ba = self.g_BA.predict(x=b_samples, batch_size=1)
        self.val_summaries.add_images("A2B_predictions", (prediction_ab + 1) / 2, epoch, dataformats="NHWC")
        self.val_summaries.add_images("B2A_predictions", (prediction_ba + 1) / 2, epoch, dataformats="NHWC")

    def update_metrics(self, metrics_dict, metrics: Dict):
        for name in metrics_dict.keys():
            metrics_dict[name].update_state(metrics[name])

    def display_metrics(self, metrics_dict, progress_bar, force: bool = False):
        """model.py:291-302.  Evaluating the running means reads device memory, i.e. waits for the steps queued so far;
        the reference does that after every batch.  Here it happens at most every `display_interval` seconds (and at
        the end of the loop), so the steps in between are enqueued back to back."""
        import time
        now = time.monotonic()
        if not force and now - self._last_display < self.display_interval:
            return
        self._last_display = now
        evaluated_metrics = {k: str(v.result())[:7] for k, v in metrics_dict.items()}
        progress_bar.set_postfix(**evaluated_metrics)

    def save_model(self):
        """model.py:304-323.  TF SavedModel cannot be written without TensorFlow: each net is saved as
        `<name>/variables.npz` (arrays in trainable_variables order); optimizer files keep the reference's
        names and `[iterations, m..., v...]` object-array layout."""
        os.makedirs(self.model_folder, exist_ok=True)
        for name, net in zip(NET_NAMES, self._nets()):
            os.makedirs(join(self.model_folder, name), exist_ok=True)
            np.savez(join(self.model_folder, name, "variables.npz"),
                     **{f"v{i:04d}": a for i, a in enumerate(net.get_weights())})
        if self._trainer is not None:
            for name, opt in zip(NET_NAMES, self._opts()):
                w = opt.get_weights()
                arr = np.empty(len(w), dtype=object)
                for i, x in enumerate(w):
                    arr[i] = x
                np.save(join(self.model_folder, f"{name}_optimizer.npy"), arr, allow_pickle=True)
        if hasattr(self, "a_samples"):
            np.save(join(self.model_folder, "a_samples.npy"), self.a_samples)
            np.save(join(self.model_folder, "b_samples.npy"), self.b_samples)

    def load_model(self):
        """model.py:325-342 against the .npz layout written by save_model."""
        for name, net in zip(NET_NAMES, self._nets()):
            z = np.load(join(self.model_folder, name, "variables.npz"))
            net.set_weights([z[k] for k in sorted(z.files)])
        self._pending_optimizer_state = {}
        for name in NET_NAMES:
            p = join(self.model_folder, f"{name}_optimizer.npy")
            if os.path.exists(p):
                self._pending_optimizer_state[name] = np.load(p, allow_pickle=True)
        for s in ("a_samples", "b_samples"):
            p = join(self.model_folder, f"{s}.npy")
            if os.path.exists(p):
                setattr(self, s, np.load(p))

    def restore_optimizers(self):
        """model.py:344-362 (`load_optimizer`): the reference applies zero gradients once to create the slots, then
        `set_weights`; here the slots are the trainer's device buffers, created zeroed.  Runs automatically when the
        native trainer is created; harmless to call again (the pending state is consumed)."""
        pending, self._pending_optimizer_state = getattr(self, "_pending_optimizer_state", {}), {}
        for i, name in enumerate(NET_NAMES):
            w = pending.get(name)
            if w is not None and len(w):
                self._optimizer_set_weights(i, list(w))

    def __del__(self):
        try:
            if self._trainer is not None:
                _lib.load().cg_trainer_destroy(self._trainer)
        except Exception:
            pass
