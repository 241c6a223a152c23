"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GaussiGAN splat renderer.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker or the timed CPU
baseline -- never as a fallback for the CUDA path.  ``gaussigan_b200`` never
imports this package.

Contents
--------
reference_port.py   float64 torch-CPU op-for-op restatement of the reference's
                    renderer (``mask/mask_gen_dynamic.py:112-143, 233-323``), its
                    fp32-raster inference flavour (``mask/infer_masks.py:186-267,
                    345-373``), the Gaussian construction / deformation tails
                    (``:378-401, 426-458``) and the direct consumers
                    (``:221-231, 785-791``).  Gradients by autograd.
closed_form.py      independent numpy float64 dual-conic closed form and
                    hand-derived backward, used
                    to cross-check reference_port.py.
tf_shim.py          a ~25-op TensorFlow-1.x name shim backed by torch, plus a
                    loader that pulls the reference's OWN function bodies out of
                    ``/root/reference`` with ``ast`` and executes them against the
                    shim.  Used (in the build container only, where
                    ``/root/reference`` exists) to generate ``tests/golden/*.npz``
                    and to pin reference_port.py to the reference source.

PARITY STATUS: the reference has no tests, golden vectors or fixtures, and its
substrate (TensorFlow 1.12 / tensorpack 0.9.0) cannot be installed here, so
parity with a *running TF build* is UNPINNED.  What IS pinned: the reference's
own Python source for the path, executed verbatim against the torch-backed shim
(``tests/golden/make_golden.py``), agrees with reference_port.py to float64
round-off, and both agree with closed_form.py and the analytic known-answer
tests.  Residual ambiguity is confined to TF primitive kernels (Eigen
``pcos/psin`` vs libm, LU pivot order, ``LinSpace`` last element), all <= 1e-6
relative -- below the 1e-5 / 1e-4 tolerances of the path.
"""
