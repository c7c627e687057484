"""CPU oracle for the GNS rollout hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the algorithm on the path that
``BASELINE.json.north_star`` names (neighbor search -> features -> GNS forward ->
integrate -> rollout loop).  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker.  ``lagrangebench_b200`` never imports it.

Parity pin status (see DESIGN.md "Oracle"):
  * space / neighbor list / features / integrate / rollout loop: PINNED against the
    reference's own golden vectors -- ``tests/case_test.py:40-206`` (3-particle periodic
    box) and ``tests/rollout_test.py:74-195`` (Lennard-Jones fixture rollout identity),
    replayed by ``tests/test_oracle_golden.py``.
  * GNS network numerics (hk.Linear / ReLU / LayerNorm / Embed / jraph.GraphNetwork /
    segment_sum): PARITY UNPINNED by the reference -- it ships no GNS test and its
    third-party dependencies (jax 0.4.29, jax-sph 0.0.3, jraph 0.0.6.dev0,
    dm-haiku 0.0.12; ``poetry.lock``) are not installable here (no network).  The
    restatement follows their published semantics and the reference call sites; the
    parameter count it implies (1 211 794 / 161 042) matches the reference's published
    "1.2M" / "161K" (``docs/pages/baselines.rst:55,62``).

All citations ``path:line`` are relative to the reference repository root.
"""

from . import space, partition, features, gns, case, rollout, metrics  # noqa: F401
