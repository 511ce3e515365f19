"""CPU oracle for the LDCBF-MPC hot path.  TEST INFRASTRUCTURE ONLY.

This package is a numpy/fp64 restatement of the reference's per-timestep algorithm
(`HumanoidNavigation/MPC/HumanoidMpc.py:380-455` and the geometry / LiDAR helpers it calls).
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product path (the `ldcbf_b200` binding and the `HumanoidNavigation`
host mirror under `humanoid-navigation-using-mpc-ldcbf_b200/`) never does: it fails loudly when the
CUDA library is missing.

Parity pinning:
* geometry (`halfplane.py`) and LiDAR (`lidar.py`) are pinned bit-for-bit against the reference's own
  functions, imported in the build container with a stub matplotlib
  (`tests/golden/make_geometry_golden.py` -> `tests/golden/geometry_golden.npz`,
  `tests/golden/lidar_golden.npz`);
* the QP restatement (`qp.py`) is pinned against the state trajectories of the reference's own IPOPT
  runs that are embedded in the report's vector PDFs (`tests/golden/make_pdf_golden.py` ->
  `tests/golden/circles_traj.npz`, `circles_delta_traj.npz`) and against the step-0 known-answer
  vector of SURVEY.md §8a;
* the solver itself (CasADi + IPOPT, un-vendored, unpinned in `requirements.txt:3`) cannot be run
  offline, so the exact optimum of the same QP is computed with a KKT-certified NNLS/LDP method.
"""
