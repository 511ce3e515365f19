"""ldcbf_b200 — Python binding of libldcbf_b200.so (hand-written sm_100a CUDA, C ABI in include/ldcbf_mpc.h).

There is no CPU fallback: importing the solver entry points without the built library raises, and every call
requires CUDA tensors.  Build with `python __graft_entry__.py build` (or `make -C .../csrc`).
"""
from .binding import (LdcbfParams, Status, abi_version, clearance_grid, default_params, half_planes, lib, lidar_cast,
                      lidar_clusters, mpc_qp,
                      mpc_step, mpc_step_packed, params_from_conf, probe_fp64, rollout, rollout_unknown)
from .batched import BatchedHumanoidMPC, BatchedUnknownEnvMPC

__all__ = ["LdcbfParams", "Status", "abi_version", "clearance_grid", "default_params", "half_planes", "lib", "lidar_cast", "lidar_clusters", "mpc_qp",
           "mpc_step", "mpc_step_packed", "params_from_conf", "probe_fp64", "rollout", "rollout_unknown", "BatchedHumanoidMPC", "BatchedUnknownEnvMPC"]
