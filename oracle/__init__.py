"""CPU oracle for the SH-voxel-grid volumetric render path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: it may be
imported by ``tests/``, by ``__graft_entry__.smoke()`` and by the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- as the checker or the CPU baseline, never as
the thing measured or shipped.  The product path (``thr3ed_atom_b200``) never imports it.

Two independent restatements of the reference algorithm live here:

* ``oracle.torch_port``  -- fp32, op-by-op PyTorch-CPU restatement (uses the same ATen
  ``grid_sample`` the reference calls, so it is also the multi-threaded CPU baseline).
* ``oracle.numpy_f64``   -- fp64 NumPy restatement with a hand-written trilinear gather and the
  analytic backward pass; arbitrates fp32 disagreements.

Parity pin: both are checked against golden vectors produced by running the *reference's own
code* (``/root/reference``, imported in the build container by ``tests/golden/make_golden.py``)
-- see ``tests/test_oracle_golden.py``.  The reference's own tests hold no numeric vectors for
this path, so these generated goldens are the pin.
"""
