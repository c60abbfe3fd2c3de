"""Input pipeline of the reference (transform/data_load.py:20-34, predict.py:20-27) on the GPU.

`normalize` keeps the reference's name and meaning (`float32(x) / 127.5 - 1`); host arrays are
normalised on the host as before (predict.py:23 feeds one image), device / uint8 batches go through
`cg_normalize_u8`.  `resize`, `random_jitter`, `apply_augmentation` and `postprocess_prediction` are the
device forms of `tf.image.resize` (bilinear, half-pixel centres), `random_jitter` (data_load.py:21-27:
resize to size+50, random crop, random horizontal flip -- ONE fused kernel) and predict.py:26-27.
The TFRecord reader (`example2image`, `create_dataset`) stays out of scope: the DVC data remote is not
reachable and TensorFlow is absent.

TensorFlow's random stream cannot be reproduced: the crop offsets and flips are drawn from a numpy
`RandomState` the caller may pass, and handed to the kernel explicitly.
"""
import ctypes

import numpy as np


def normalize(tensor) -> np.ndarray:
    """data_load.py:31-34 for host arrays (any integer / float dtype): float32(x) / 127.5 - 1."""
    image = np.asarray(tensor, dtype=np.float32)
    return (image / np.float32(127.5)) - np.float32(1.0)


def _env():
    from .. import _lib
    from ..runtime import DeviceTensor, _ptr, _require_cuda, _stream_ptr
    torch = _require_cuda()
    return torch, _lib, DeviceTensor, _ptr, _stream_ptr


def _as_cuda(x, torch, dtype):
    from ..runtime import DeviceTensor
    if isinstance(x, DeviceTensor):
        x = x.torch
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
    if x.dtype != dtype:
        x = x.to(dtype)
    return x.cuda().contiguous()


def normalize_device(images_u8):
    """normalize (data_load.py:31-34) of a uint8 batch on the GPU -> float32 DeviceTensor of the same shape."""
    torch, _lib, DeviceTensor, _ptr, _stream_ptr = _env()
    x = _as_cuda(images_u8, torch, torch.uint8)
    y = torch.empty(x.shape, dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().cg_normalize_u8(_ptr(x), _ptr(y), x.numel(), _stream_ptr(torch)), "cg_normalize_u8")
    return DeviceTensor(y)


def postprocess_prediction(prediction):
    """predict.py:26-27: np.array((prediction[0] + 1) * 127.5, np.uint8) -- here for the whole batch, on the GPU.
    Returns a uint8 numpy array [N, H, W, C]; index [0] for the reference's single-image result."""
    torch, _lib, DeviceTensor, _ptr, _stream_ptr = _env()
    x = _as_cuda(prediction, torch, torch.float32)
    y = torch.empty(x.shape, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().cg_postprocess_u8(_ptr(x), _ptr(y), x.numel(), _stream_ptr(torch)), "cg_postprocess_u8")
    return y.cpu().numpy()


def resize(images, size):
    """tf.image.resize(images, size) (data_load.py:23,41): bilinear, half-pixel centres, no antialiasing; float32 NHWC."""
    torch, _lib, DeviceTensor, _ptr, _stream_ptr = _env()
    x = _as_cuda(images, torch, torch.float32)
    squeeze = x.dim() == 3
    if squeeze:
        x = x[None]
    N, H, W, C = x.shape
    Ho, Wo = int(size[0]), int(size[1])
    y = torch.empty((N, Ho, Wo, C), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().cg_resize_bilinear(_ptr(x), N, H, W, C, _ptr(y), Ho, Wo, _stream_ptr(torch)),
               "cg_resize_bilinear")
    return DeviceTensor(y[0] if squeeze else y)


def random_jitter(images, image_size: int, rng=None, return_draws: bool = False):
    """random_jitter (data_load.py:21-27) for a batch: resize to [image_size + 50]^2, random crop to
    [image_size]^2, random horizontal flip -- one kernel, the resized image is never materialised."""
    torch, _lib, DeviceTensor, _ptr, _stream_ptr = _env()
    rng = rng if rng is not None else np.random.RandomState()
    x = _as_cuda(images, torch, torch.float32)
    squeeze = x.dim() == 3
    if squeeze:
        x = x[None]
    N, H, W, C = x.shape
    big = image_size + 50
    oy = rng.randint(0, big - image_size + 1, size=N).astype(np.int32)
    ox = rng.randint(0, big - image_size + 1, size=N).astype(np.int32)
    flip = (rng.uniform(size=N) < 0.5).astype(np.int32)
    d = [torch.from_numpy(a).cuda() for a in (oy, ox, flip)]
    y = torch.empty((N, image_size, image_size, C), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().cg_resize_crop_flip(_ptr(x), N, H, W, C, big, big, _ptr(y), image_size, image_size,
                                               _ptr(d[0]), _ptr(d[1]), _ptr(d[2]), _stream_ptr(torch)),
               "cg_resize_crop_flip")
    out = DeviceTensor(y[0] if squeeze else y)
    return (out, (oy, ox, flip)) if return_draws else out


def apply_augmentation(dataset, image_size: int, rng=None):
    """data_load.py:20-29 over a python iterable of HWC images: yields jittered float32 HWC numpy samples."""
    rng = rng if rng is not None else np.random.RandomState()
    for image in dataset:
        yield random_jitter(image, image_size, rng).numpy()
