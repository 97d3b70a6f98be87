"""CPU oracle for the BORE-MLP hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy, the arithmetic that the reference
(ltiao/bore v1.5.0) delegates to TensorFlow-Keras 2.5.0, and drives the *installed*
SciPy L-BFGS-B for the optimiser half.  It exists so the CUDA path can be checked; it
is never imported by ``bore_b200`` (the product).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.

Parity status
-------------
* L-BFGS-B half: **pinned** to the live SciPy (1.18.1 here; a C translation of the
  same L-BFGS-B v3.0 that the reference's pinned scipy==1.7.0 wraps in Fortran).
* Keras half (Dense / activations / BCE / Adam / fit loop): **parity unpinned** --
  TensorFlow is not installable in this image, and the reference's own tests hold no
  numeric known-answers for this path (tests/test_models.py:12-50 is a property test).
  The restatement follows TF 2.5.0 semantics from knowledge of that release; value and
  gradient are cross-checked against torch-CPU autograd (tests/test_oracle.py).
* Host helpers (math.py, optimizers/utils.py, data.py): pinned against the reference
  modules themselves, which import fine here (tests/golden/make_golden.py).
* Data step (data_step.py) and SVGD batch argmax (svgd.py): **pinned** -- bore.data and
  bore.optimizers.svgd import here (numpy / scipy / sklearn only); the fixtures
  tests/golden/{data_step,svgd}_golden.npz are outputs of the reference code itself and the
  restatements reproduce them exactly.
"""
