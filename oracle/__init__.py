"""CPU oracle for the posg-b200 hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package; the product
package ``posggym_baselines_b200`` never does (tests/test_no_oracle_in_product.py
greps for it).

What it restates
----------------
The arithmetic of the hot path lives in the third-party dependency ``posggym``
(requirement ``posggym[agents] >=0.5.0``, /root/reference/pyproject.toml:36; no pin,
no lock file, not vendored, not installable here: no network).  The modules
below restate the published algorithm of

  posggym/envs/grid_world/core.py                 -> oracle/grid.py
  posggym/envs/grid_world/driving.py   (Driving-v1)            -> oracle/driving.py
  posggym/envs/grid_world/predator_prey.py (PredatorPrey-v0)   -> oracle/predator_prey.py
  posggym/envs/grid_world/level_based_foraging.py (LBF-v3)     -> oracle/lbf.py
  posggym/model.py, posggym/utils/history.py, gymnasium.spaces  -> oracle/shims/
  posggym/vector/sync_vector_env.py + wrappers (StackEnv,
  RecordEpisodeStatistics, FlattenObservations)                 -> oracle/vec_env.py

with every random draw hoisted into an explicit ``noise`` argument so that
"identical (state, joint-action, noise)" is well defined
(``step_with_noise``); the POSGModel-shaped methods (``step``,
``sample_initial_state`` ...) draw that noise from ``model.rng``.

PARITY STATUS: **unpinned against the upstream source** (it is not in the
container), pinned against every fixture the reference tree holds for this
path -- see tests/test_oracle_fixtures.py and DESIGN.md "Oracle pinning":
flattened-observation layout (Adam second-moment masks of the shipped BR
checkpoints), the 14x14RoundAbout start/destination distance histogram,
the float32 reward residuals of 10 671 successful episodes, the return lattices
and the uniform-random joint-play statistics of ``div_results.csv``.

The unmodified reference planner (``/root/reference/posggym_baselines/planning``)
runs on these models through ``oracle.shims.install()`` -- that is the
planner-level oracle used for the statistical parity tests.
"""
