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
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
