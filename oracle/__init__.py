"""CPU oracle for the lagrangian-microbes per-timestep hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the timed CPU
baseline.  The product (``lagrangian_microbes_b200``) never imports this package
and raises if its CUDA library is missing.

Parity pins:
  * pair search  -- pinned: ``scipy.spatial.cKDTree.query_pairs`` is the very
    library call the reference makes (interaction_simulator.py:93,98); the
    restated brute-force predicate is checked against it.
  * RPS rule     -- pinned: the restatement is checked against golden vectors
    produced by the UNMODIFIED reference function
    (/root/reference/interactions.py:13-40) -- tests/golden/make_golden.py.
  * RK4 advection -- PARITY UNPINNED: the arithmetic lives in parcels
    2.0.0beta2 (environment.yml:80), which is neither vendored nor installable
    here and the reference has no test touching it.  ``oracle/rk4.py`` restates
    the published algorithm; it is anchored only by analytic known-answer tests.
  * delta-packed record (``oracle/record.py``) -- PARITY UNPINNED by construction: the
    format is this framework's own (the reference stores plain float32 columns);
    its contract is the bit-exact round trip back to those columns.
"""
