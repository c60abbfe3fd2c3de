"""cyclegan_cat_b200 -- B200-native (sm_100a) CycleGAN training step behind the
interface of dogeplusplus/cyclegan-cat.

Layout:
  cyclegan/          same module / function names as the reference package `cyclegan/`
  model_processing/  yaml2namespace / namespace2yaml (reference: model_processing/load_model.py)
  transform/         normalize (reference: transform/data_load.py:31-34)
  ir.py, runtime.py  layer-graph IR and the host objects that call the C-ABI
  csrc/              hand-written CUDA + the C-ABI (include/cyclegan_b200.h)
  lib/               libcyclegan_b200.so, built in-tree by build.py (git-ignored)

Importing this package never imports `oracle/`.
"""
from . import ir  # noqa: F401

__all__ = ["ir"]
