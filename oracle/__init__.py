"""CPU oracle for the keras_nerf per-ray hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch-CPU fp32 / numpy) restatement of the arithmetic
of naufalso/keras_nerf's hot path.  It is the *checker* for the CUDA product
in ``keras_nerf_b200/``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product path never imports it and fails loudly when ``libknerf.so`` is
missing.

Parity pin status
-----------------
The reference is Python on TensorFlow >= 2.9 (``requirements.txt:2``, no upper
pin, no lock file).  TensorFlow is not installed in the build image nor on the
GPU box and cannot be installed (no network), so the reference's own kernels
cannot be executed.  The oracle is pinned in two ways:

1. the reference's only known-answer test (``tests/data/test_utils.py:5-10``,
   focal from fov) and the range/determinism asserts of
   ``tests/data/test_rays.py:50-87`` are re-run against the oracle;
2. ``oracle/tfshim`` is a small torch-CPU stand-in for the ~50 ``tf.*`` symbols
   the hot path touches.  With it on ``sys.path`` the reference's OWN Python
   source (``/root/reference/keras_nerf/...``: ``RaysGenerator``, ``NeRFUtils``,
   ``NeRFMLP``, ``NeRF.train_step`` ...) is imported and executed unmodified,
   and its outputs are committed as ``tests/golden/*.npz`` by
   ``tests/golden/make_golden.py``.  The oracle is checked against those
   fixtures, so every line of reference *Python* (op order, quirks, shapes) is
   pinned; what remains un-pinned is the arithmetic inside the TF ops
   themselves (Eigen/cuBLAS rounding, reduction order), which the shim
   restates from TF's documented semantics.  Where those semantics are device
   dependent (out-of-range ``tf.gather``: error on CPU, zero on GPU) the
   choice is explicit (``oob_mode``) and documented in DESIGN.md.

So: "parity pinned to the reference's Python source executed over a TF
stand-in; TF op internals unpinned (TensorFlow unavailable offline)".
"""
from .nerf_oracle import *  # noqa: F401,F403
