"""Diagnostic: run-to-run spread of the metrics of 5 training steps, graph replay vs eager and eager vs eager
(fp32 check mode, the nets of tests/test_gpu_api.py::test_cuda_graph_replay_matches_eager_steps)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import common as C                                   # noqa: E402
from tests.test_gpu_api import _dataset, _gan                   # noqa: E402


def run(disable, mode):
    os.environ["CG_DISABLE_GRAPH"] = disable
    gan = _gan("/tmp/cg_diag", mode=mode)
    for i, n in enumerate((gan.g_AB, gan.g_BA, gan.d_A, gan.d_B)):
        n.initialize(7 + i)
    data = _dataset(4, size=64, seed=3)
    a, b = np.stack([t[0] for t in data[:2]]), np.stack([t[1] for t in data[:2]])
    return [float(gan.train_step(a, b)["gAB_loss"]) for _ in range(5)]


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    runs = [("graph", run("0", mode)), ("eager", run("1", mode)), ("eager", run("1", mode)), ("graph", run("0", mode))]
    base = runs[1][1]
    for name, r in runs:
        print(name, " ".join(f"{x:.7f}" for x in r), "| rel diff vs first eager:",
              " ".join(f"{abs(x - y) / abs(y):.1e}" for x, y in zip(r, base)))
