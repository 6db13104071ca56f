"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the tornadox hot path (the parity oracle).

Nothing under ``tornadox_b200/`` may import from here.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use this package, and only as the checker / the timed CPU baseline.

* ``oracle.ref_np``   NumPy fp64 restatement, op-for-op with the reference (file:line cited
                      in every docstring).  ``numpy.linalg.qr/cholesky`` dispatch to the same
                      LAPACK ``geqrf``/``potrf`` routines JAX-CPU uses.
* ``oracle.jaxshim``  a NumPy-backed stand-in for the handful of ``jax`` entry points the
                      reference imports, so that the UNMODIFIED reference sources under
                      ``/root/reference`` can be executed in the build container (JAX itself is
                      not installable here).  Used by ``oracle/make_golden.py`` to generate
                      ``tests/golden/*.npz`` and to pin ``ref_np`` against the reference's code.

Parity pin status: see DESIGN.md section "Oracle".
"""
