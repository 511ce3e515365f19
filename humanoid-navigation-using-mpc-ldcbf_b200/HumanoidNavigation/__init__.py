"""Host-side mirror of the reference's `HumanoidNavigation` package for the LDCBF-MPC hot path.

Same module paths, class names, constructor signatures and return shapes as the reference, so a caller written
against `HumanoidNavigation.MPC.HumanoidMpc.HumanoidMPC(...).run_simulation(...)` runs unchanged; the per-step
arithmetic is executed by hand-written sm_100a kernels through `ldcbf_b200` (no CasADi, no CPU fallback).
Plotting / animation (matplotlib) is out of scope: `run_simulation` returns `animator = None`.
"""
