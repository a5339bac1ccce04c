"""CPU oracle for the auv_particle_filter measurement cycle.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import, link or execute it, and
there only as the checker / the CPU arm being timed.  The product path
(``smarc_navigation_b200``) never imports this package and fails loudly when
its CUDA library is missing.

Parity status:

* propagate / roughen / GPS weight / normalise / resamplers / resample
  bookkeeping / estimate: restated from the reference
  (``auv_particle_filter/scripts/{auv_particle,resampling,auv_pf}.py``) and
  PINNED against golden vectors produced by importing the unmodified reference
  modules (``tests/golden/make_golden.py``; the reference has no tests or
  fixtures of its own).
* ``tf.transformations`` (ROS ``geometry`` stack, un-vendored, version unpinned
  by ``auv_particle_filter/package.xml:57,71``) is restated in ``tf_stub.py``
  from its published algorithm and cross-checked against SciPy's ``Rotation``.
* MBES raycast + beam-range likelihood: the reference contains NO
  implementation -- **parity unpinned**.  The oracle here is
  the build's own specification (``raycast_spec`` in ``c/auvpf_oracle.c``),
  anchored by an independent brute-force Moller-Trumbore check in
  ``raycast_bruteforce.py``.
"""
