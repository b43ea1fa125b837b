"""voltool_b200 -- B200 (sm_100a) implementation of voltool's data-parallel hot path.

The product is the C-ABI shared library ``libvoltool_gpu.so`` (``include/voltool_gpu.h``), built from
``voltool_b200/csrc`` by ``__graft_entry__.build()`` / ``make -C voltool_b200/csrc``.  This package is
the thin Python face of that ABI, used by the tests and the benchmark; it mirrors the reference's
operator names (VolumetricData methods, cc_threaded, add/subtract/..., fit) one to one.

There is no CPU fallback: importing works anywhere (so host-only logic can be tested), but every
compute entry point raises ``VoltoolGpuError`` unless the library is built and a CUDA device
initialises.  Nothing in this package imports, links or calls ``oracle/``.
"""
from .api import (Grid, Cache, Pose, FitResult, VolMap, VoltoolGpuError, lib, lib_path, init, shutdown, sync,
                  launch_count, device_props, cc_threaded, add, subtract, multiply, average, binary, sim, sim_policy,
                  calc_cc, sim_cc, fit, FitContext, rotate_replay, pose_apply, measure_center, fit_schedule,
                  init_from_intersection, init_from_union, profile_enable, profile_reset, profile_get, write_dx,
                  read_dx, write_mrc, read_mrc, cc_host, binary_host, compact_selection, sim_sel, sim_cc_sel, fit_sel,
                  nccl_unique_id, nccl_init, nccl_shutdown, best_pose_device, nccl_allgather, cc_maps, sim_cc_maps,
                  density_rotate, density_reset)

__all__ = ["Grid", "Cache", "Pose", "FitResult", "VolMap", "VoltoolGpuError", "lib", "lib_path", "init", "shutdown",
           "sync", "launch_count", "device_props", "cc_threaded", "add", "subtract", "multiply", "average", "binary",
           "sim", "sim_policy", "calc_cc", "sim_cc", "fit", "FitContext", "rotate_replay", "pose_apply",
           "measure_center", "fit_schedule", "init_from_intersection", "init_from_union", "profile_enable",
           "profile_reset", "profile_get", "write_dx", "read_dx", "write_mrc", "read_mrc", "cc_host", "binary_host",
           "compact_selection", "sim_sel", "sim_cc_sel", "fit_sel", "nccl_unique_id", "nccl_init", "nccl_shutdown",
           "best_pose_device", "nccl_allgather", "cc_maps", "sim_cc_maps", "density_rotate", "density_reset"]
