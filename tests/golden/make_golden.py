"""Regenerate the frozen oracle fixtures:  python tests/golden/make_golden.py

TensorFlow cannot be imported in this environment, so these vectors come from the torch-CPU
restatement in oracle/ (parity unpinned) -- they guard the oracle against drift and give the GPU
tests a fixed target.  The reflect-pad vector is the reference's own (unittests/test_resnet.py:31-47).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.train import OracleCycleGan, synthetic_batch  # noqa: E402
from tests import common as C  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    o = OracleCycleGan(C.SMALL_UNET, C.SMALL_SIMPLE)
    a, b = synthetic_batch(1, 32)
    for step in range(2):
        for k, v in o.train_step(a, b).items():
            out[f"step{step}_{k}"] = np.float32(v)
    out["g_AB_var0"] = o.g_AB.variables[0].detach().numpy()
    np.savez(os.path.join(HERE, "oracle_c1_small.npz"), **out)

    # a second fixture the GPU tests compare against directly: resnet + simple D, metrics and a few gradients
    o = OracleCycleGan(C.SMALL_RESNET, C.SMALL_SIMPLE)
    a, b = synthetic_batch(2, 32)
    metrics, grads, imgs = o.gradients(a, b)
    out = {f"metric_{k}": np.float32(v) for k, v in metrics.items()}
    for net in ("g_AB", "g_BA", "d_A", "d_B"):
        for i in (0, 2, len(grads[net]) - 2):
            out[f"grad_{net}_{i}"] = grads[net][i].numpy()
    out["fake_b"] = imgs["fake_b"].numpy()
    np.savez(os.path.join(HERE, "oracle_resnet_small.npz"), **out)
    np.savez(os.path.join(HERE, "reflect_pad_reference.npz"),
             x=np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2]])[np.newaxis, ..., np.newaxis],
             expected=np.array([[1, 1, 1, 1, 1], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1], [2, 2, 2, 2, 2],
                                [1, 1, 1, 1, 1]])[np.newaxis, ..., np.newaxis])
    # optional config paths: BatchNormalization + Dropout nets under RMSprop / AdaBelief, two steps + an inference step;
    # the dropout mask bits and a resize / jitter sample of the input pipeline
    from oracle import tf_ops as T
    o = OracleCycleGan(C.BN_DROP_UNET, C.BN_SIMPLE, g_opt=dict(name="adabelief", learning_rate=2e-4),
                       d_opt=dict(name="rmsprop", learning_rate=2e-4))
    for i, n in enumerate(("g_AB", "g_BA", "d_A", "d_B")):
        getattr(o, n).drop_seed = 100 + i
    a, b = synthetic_batch(2, 32)
    out = {}
    for step in range(2):
        for k, v in o.train_step(a, b).items():
            out[f"step{step}_{k}"] = np.float32(v)
    for k, v in o.validate_step(a, b).items():
        out[f"val_{k}"] = np.float32(v)
    out["g_AB_moving_mean0"] = o.g_AB.state[0].numpy()
    out["g_AB_moving_var0"] = o.g_AB.state[1].numpy()
    out["d_A_moving_var1"] = o.d_A.state[3].numpy()
    out["g_AB_var0"] = o.g_AB.variables[0].detach().numpy()
    out["d_A_var0"] = o.d_A.variables[0].detach().numpy()
    out["dropout_mask"] = T.dropout_mask(100, 1, 2, 3, 4096, 0.5).astype(np.uint8)
    x = np.random.RandomState(4).uniform(-1, 1, (2, 20, 24, 3)).astype(np.float32)
    out["resize_in"] = x
    out["resize_out"] = T.resize_bilinear(x, 33, 17)
    out["jitter_out"] = T.random_jitter(x, 16, np.array([5, 50]), np.array([0, 31]), np.array([1, 0]))
    np.savez(os.path.join(HERE, "oracle_options.npz"), **out)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
