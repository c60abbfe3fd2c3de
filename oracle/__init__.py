"""CPU oracle for the CycleGAN hot path -- TEST INFRASTRUCTURE ONLY.

This package is a torch-CPU restatement of the reference's TensorFlow/Keras
training step (``/root/reference/cyclegan/*.py``).  It exists to *check* the
CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package (``cyclegan_cat_b200``) never imports anything from here
and fails loudly when its CUDA library is missing.

PARITY UNPINNED: TensorFlow 2.7 / tensorflow-addons 0.15 are not installable
in this environment (no network), and the reference's own tests hold exactly
one known-answer vector (reflection padding, ``unittests/test_resnet.py:31-47``)
plus output shapes.  The oracle is pinned against those; for conv / instance
norm / losses / gradients / Adam it restates the published TF semantics
 and is self-consistency-checked only.
"""
