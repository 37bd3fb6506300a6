"""CPU oracle for the gpitch variational-GP hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker (or as the timed CPU baseline), never as the thing shipped.  ``gpitch_b200`` never
imports this package.

What it is: an op-for-op fp64 restatement (torch-CPU, so autograd supplies reference gradients the
way ``tf.gradients`` does in the reference) of

* the on-disk reference arithmetic -- ``gpitch/matern12_spectral_mixture.py``,
  ``gpitch/sgpr_ss.py``, ``gpitch/pdgp.py``, ``gpitch/likelihoods.py``, ``gpitch/methods.py``,
  ``gpitch/window_overlap.py`` (each function cites the file:line it follows), and
* the GPflow 0.5 / TensorFlow 1.2.1 semantics those files call (``Stationary``, ``Matern32``,
  ``Add``, ``conditional``, ``gauss_kl``, ``SGPR.build_predict``, transforms).  GPflow 0.5 is a
  third-party dependency that is NOT on disk (pinned only by the reference notebooks' cell-2
  output ``gpflow 0.5`` / ``tf 1.2.1``); those formulas are restated from the published algorithm
   and are marked ``[GPflow-0.5, recalled]`` in ``gpflow_ref.py``.

Pinning status
--------------
* PINNED against the reference's own source: everything on disk.  ``oracle/refshim`` executes the
  UNMODIFIED reference files from ``/root/reference`` under a TensorFlow->torch API shim
  (``oracle/make_golden.py``), and the resulting vectors are committed in ``tests/golden/``;
  ``tests/test_oracle_golden.py`` checks this restatement against them.  Window indexing is pinned
  bit-exactly the same way, and the two reproducible known answers of the reference notebooks
  (109 inducing points, f0 = 261.6255653005986 Hz) are tests as well.
* PARITY UNPINNED for the GPflow-0.5 pieces (``conditional``, ``gauss_kl``, ``Stationary``
  distance, ``SGPR.build_predict``, ``positive`` transform constant 1e-6, jitter 1e-6): the
  reference holds no golden vectors or tests for the hot path and GPflow 0.5 cannot be installed
  here.  They are cross-checked through mathematical invariants instead (SVGP bound at the optimal
  q == collapsed SGPR bound, mpmath evaluation, finite differences) -- see tests/test_oracle_*.py.
"""
