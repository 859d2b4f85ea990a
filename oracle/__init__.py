"""CPU oracle for the Lovász-Softmax + confusion-matrix mIoU hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU arm.  The shipped package
(``miccai2021_cataract_semantic_segmentation_b200``) never imports this module
and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``tests/golden/cases.npz`` holds input/output vectors
produced by importing the unmodified reference from ``/root/reference`` in the
build container (``tests/golden/make_golden.py``, committed);
``tests/test_oracle_golden.py`` checks both restatements below against every
one of them.

* ``oracle.port``  — torch-CPU restatement (same ATen ops and op order as the
  reference, ``stable=True`` sort), multi-threaded; also the timed CPU arm.
* ``oracle.cref``  — ctypes binding of ``oracle/lovasz_cm_ref.c``, a plain-C
  scalar restatement with the closed-form backward (independent second opinion
  for the math the CUDA kernels implement).
"""
