"""CPU oracle for the Dual-DMP training hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package, and only as the checker or the timed CPU baseline. The product (``dual_dmp_b200``) never
imports it and has no CPU fallback.

What is restated, and how it is pinned:

* ``oracle.loss_ref`` / ``oracle.models_ref`` / ``oracle.mesh_ref`` restate the reference's own
  ``util/loss.py``, ``util/models.py`` and ``util/mesh.py``.  PINNED: ``oracle/make_golden.py`` imports the real
  reference modules from ``/root/reference`` (with a ``pymeshlab`` stub) in the build container, runs them on
  seeded inputs and commits inputs + outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the
  restatement against those vectors.
* ``oracle.gcn_ref`` / ``oracle.networks_ref`` restate ``torch_geometric==2.2.0`` ``GCNConv`` (+``gcn_norm``,
  ``add_remaining_self_loops``, glorot init) and ``torch_scatter==2.1.0`` ``scatter_add`` as used by
  ``util/networks.py:15-26,51-62``.  Those packages are pinned by the reference (``requirements.txt:15,19``) but
  are neither vendored in ``/root/reference`` nor installable here, and the reference has no tests or golden
  vectors at that boundary.  PARITY UNPINNED for this part: it is anchored on the published
  algorithm  Y = D^-1/2 (A+I) D^-1/2 X W^T + b, cross-checked against an independent dense float64 evaluation
  and hand-computed tiny graphs (``tests/test_oracle_gcn.py``).
"""
