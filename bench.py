#!/usr/bin/env python
"""bench.py -- voltool fit pose-evaluation throughput on B200 (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (torchrun for N > 1)
  python bench.py --impl reference --gpus N ...             the reference's CPU path on the host cores

One step = one batch of rigid poses (rotation x translation candidates of the cfg2 grid search)
evaluated for a 50k-atom structure against a resident 256^3 target map: transform -> Gaussian
splat on the pose's own grid -> thresholded CC (threshold 0.1) -> per-shard best.  Each rank
evaluates its own slice of the pose list (weak scaling, no data-path collective); the per-shard
best (cc, pose index) is merged with one small all_gather per step.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each field means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PLANTED = (58.0, 29.0, 315.0)     # 48/24/312 on the 24-degree lattice + 8/4/0 + 2/1/3 refinements
RESOLUTION = 5.0
FLOP_PER_PAIR = 7.0               # 1 exp + ~6 FP32 flops per (atom, voxel) pair in the cutoff box


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line):
    NVML from a thread every 5 ms (the region can be shorter than one nvidia-smi period); nvidia-smi
    -lms as the fallback when NVML is not importable"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, uuid=None):
        self.index, self.uuid = index, uuid
        self.proc, self.lines = None, []
        self.nv, self.handle, self.stop_flag, self.thread = None, None, False, None
        self.sm, self.mx, self.mask = [], [], 0

    def _nvml_open(self):
        try:
            if os.environ.get("BENCH_NO_NVML"):
                return False
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                for u in (self.uuid, "GPU-" + self.uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(u.encode() if isinstance(u, str) else u)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nv, self.handle = pynvml, h
            return True
        except Exception:
            return False

    def _nvml_loop(self):
        nv, h = self.nv, self.handle
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self._nvml_open():
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            reasons = sorted(k for k, bit in self.BITS.items() if self.mask & bit)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                    "sm_max_mhz": max(self.mx) if self.mx else None, "samples": len(self.sm), "reasons": reasons,
                    "source": "nvml, 5 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def pose_list():
    """10 000 rotations (25 x 20 x 20 Euler lattice) x 27 translations (+-1 voxel) = 270 000 poses"""
    rx = np.arange(25) * (360.0 / 25)
    ry = np.arange(20) * (360.0 / 20)
    rz = np.arange(20) * (360.0 / 20)
    rot = np.stack(np.meshgrid(rx, ry, rz, indexing="ij"), -1).reshape(-1, 3)
    sh = np.stack(np.meshgrid([-1.0, 0.0, 1.0], [-1.0, 0.0, 1.0], [-1.0, 0.0, 1.0], indexing="ij"), -1).reshape(-1, 3)
    poses = np.concatenate([np.repeat(rot, len(sh), 0), np.tile(sh, (len(rot), 1))], 1).astype(np.float32)
    # interleave so that every slice mixes rotations (pose cost varies with the rotated bounding box)
    perm = np.random.default_rng(2003).permutation(len(poses))
    return np.ascontiguousarray(poses[perm])


def splat_pairs(radii, radscale, gspacing, gausslim=4.0):
    """algorithmic (atom, voxel) pairs of one splat: sum over atoms of prod over axes (2*floor(cutoff/gs)+1)"""
    b = 2 * np.floor(gausslim * radii.astype(np.float64) * radscale / gspacing) + 1
    return float(np.sum(b ** 3))


def splat_traffic(args, poses_per_launch):
    """dram__bytes_read + dram__bytes_write of one splat_list_kernel launch, as MEASURED by the committed ncu --set full
    capture of this build (profiles/r2l_splat_traffic.json is written from the capture's summary); only quoted for the
    configuration that capture ran (cfg2, same poses per launch), else null"""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2l_splat_traffic.json")
    try:
        t = json.load(open(path))
    except Exception:
        return None, "no ncu capture of this build under profiles/"
    if args.atoms != 50000 or args.map != 256 or abs(poses_per_launch - t["poses_per_launch"]) > 0.5:
        return None, f"{t['source']} is of cfg2 at {t['poses_per_launch']} poses per launch, not this configuration"
    return t["traffic"], t["source"]


def build_cfg2(vt, atoms, n):
    """BASELINE configs[1] on the device: `atoms` atoms in a ball, target = 1 A density of the structure in the
    PLANTED pose, zero-padded to n^3, + N(0, 0.02 max) noise; the structure is then COM-aligned to the map as
    fit() does (TclVoltool.C:789-803).  Returns pos, radii, mass, com, resident target map."""
    from voltool_b200.synth import make_structure, add_noise
    pos, radii, mass = make_structure(atoms, 52.0 * (atoms / 50000.0) ** (1 / 3), seed=2001, center=(0.5 * (n - 1),) * 3)
    com0 = vt.measure_center(pos, mass)
    planted = vt.pose_apply(pos, com0, PLANTED)
    radscale, gsp, quality = vt.sim_policy(RESOLUTION)
    tmap = vt.sim(planted, radii, quality, radscale, 1.0)      # 1 A sampling of the planted pose
    g = tmap.grid
    padm = [(n - s) // 2 for s in (g.xsize, g.ysize, g.zsize)]
    padp = [n - s - m for s, m in zip((g.xsize, g.ysize, g.zsize), padm)]
    tmap.pad(padm[0], padp[0], padm[1], padp[1], padm[2], padp[2])
    tmap.upload(add_noise(tmap.data, 0.02, seed=2002))
    tg = tmap.grid
    assert (tg.xsize, tg.ysize, tg.zsize) == (n, n, n)
    dencom = tmap.vol_com()
    pos = (pos + (dencom - com0)[None, :]).astype(np.float32)   # fit's COM alignment (TclVoltool.C:800)
    com = vt.measure_center(pos, mass)
    return pos, radii, mass, com, tmap


# ================================================================================================ ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    import voltool_b200 as vt
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner with a plain printf when the first communicator comes up: send
        # stdout to stderr while that happens, so that stdout carries only the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    vt.init(local)
    if world > 1:
        from voltool_b200 import dist as vdist
        vdist.init_nccl_from_torch()      # the LIBRARY's communicator: the path's one collective is issued by the library
    props = vt.device_props()
    stream = torch.cuda.ExternalStream(vt.lib().vt_stream(), device=torch.device("cuda", local))
    peaks = load_peaks()

    # ---- workload (every rank holds the full target map and atom set) ----
    n = args.map
    pos, radii, mass, com, tmap = build_cfg2(vt, args.atoms, n)
    radscale, gsp, quality = vt.sim_policy(RESOLUTION)
    ctx = vt.FitContext(pos, radii, tmap, RESOLUTION)

    poses = pose_list()
    P = args.poses
    W, K = args.warmup, args.steps

    def slice_for(step):
        base = ((step * world + rank) * P) % len(poses)
        idx = (base + np.arange(P)) % len(poses)
        return poses[idx]

    d_cc = torch.empty(P, dtype=torch.float32, device="cuda")
    best_local = torch.empty(2, dtype=torch.float32, device="cuda")
    gathered = torch.empty(2 * world, dtype=torch.float32, device="cuda")
    dev_poses = [torch.from_numpy(slice_for(s)).cuda() for s in range(W + K)]
    torch.cuda.synchronize()

    def step_resident(s):
        # all three on the library stream: pose pipeline, per-shard best (cc, pose index), ONE small ncclAllGather
        ctx.eval_poses_device(com, dev_poses[s].data_ptr(), P, d_cc.data_ptr())
        vt.best_pose_device(d_cc.data_ptr(), P, ((s * world + rank) * P) % len(poses), best_local.data_ptr())
        vt.nccl_allgather(best_local.data_ptr(), gathered.data_ptr(), 2)

    def barrier():
        vt.sync()                      # the library stream first: its collective never overlaps torch's barrier
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, first, count):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
        for s in range(first, first + count):
            fn(s)
        with torch.cuda.stream(stream):
            e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for s in range(W):
        step_resident(s)
    launches0 = vt.launch_count()
    vt.profile_reset()
    vt.profile_enable(True)
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, W, K)
    clocks = sampler.stop() if rank == 0 else None
    launches = vt.launch_count() - launches0           # every kernel of the step is the library's (the gather is NCCL's)
    prof = {k: vt.profile_get(k) for k in ("transform", "setup", "splat", "pose_cc")}
    vt.profile_enable(False)
    value = world * P * K / (ms_total * 1e-3)
    best = gathered.cpu().numpy().reshape(world, 2)

    # ---- e2e: the C-ABI call with HOST buffers (coordinates + pose list in, cc table out) ----
    host_slices = [slice_for(s) for s in range(W + K)]
    pinned = torch.from_numpy(pos.copy()).pin_memory()

    mine = torch.empty(2, dtype=torch.float32, device="cuda")
    mine_host = torch.empty(2, dtype=torch.float32).pin_memory()

    def step_e2e(s):
        ctx.set_origin(pinned.numpy())                       # H2D: 3*natoms floats (+ group spheres)
        cc = ctx.eval_poses(com, host_slices[s])             # H2D: pose list; D2H: cc per pose
        i = int(np.nanargmax(cc))
        mine_host[0], mine_host[1] = float(cc[i]), float(((s * world + rank) * P) % len(poses) + i)
        with torch.cuda.stream(stream):
            mine.copy_(mine_host, non_blocking=True)
        vt.nccl_allgather(mine.data_ptr(), gathered.data_ptr(), 2)
        vt.sync()
        return float(gathered[0].item())                     # D2H read of the merged result

    step_e2e(0)
    t0 = time.perf_counter()
    ms_e2e = timed(step_e2e, W, K)
    wall_e2e = time.perf_counter() - t0
    e2e_value = world * P * K / (ms_e2e * 1e-3)
    h2d = args.atoms * 16 + (args.atoms + 31) // 32 * 16 + P * 36
    d2h = P * 4 + 4

    # ---- strong scaling of ONE `voltool fit` (the reference's own 5-stage schedule, TclVoltool.C:895-909): every rank
    # calls vt_fit on the same inputs; the library shards each rotate() table (rank r: candidates r, r + N, ...), merges it
    # with one ncclAllGather on device buffers per stage and replays the skip rule identically on every rank
    fit_sharded = None
    if not args.no_extras:
        vt.fit(pos, radii, mass, tmap, RESOLUTION)           # warm-up: grows the memory pool by the workspace
        reps, times, dev_ms = 3, [], []
        for _ in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record()
            t0 = time.perf_counter()
            p_fit, fres = vt.fit(pos, radii, mass, tmap, RESOLUTION)
            with torch.cuda.stream(stream):
                e1.record()
            vt.sync()
            times.append(time.perf_counter() - t0)
            dev_ms.append(e0.elapsed_time(e1))
        tt = torch.tensor([min(times), min(dev_ms)], device="cuda", dtype=torch.float64)
        sig = torch.tensor([float(np.float64(p_fit.astype(np.float64).sum())), float(fres.best_cc)], device="cuda", dtype=torch.float64)
        sig_all = [torch.zeros_like(sig) for _ in range(world)]
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_gather(sig_all, sig)
        else:
            sig_all = [sig]
        same = all(bool(torch.equal(x, sig_all[0])) for x in sig_all)
        fit_sharded = {"seconds": float(tt[0].item()), "device_ms": float(tt[1].item()), "ranks": world,
                       "candidates_computed": int(fres.n_computed), "candidates_visited_by_reference_loop": int(fres.n_evaluated),
                       "best_cc": float(fres.best_cc), "stage_best": list(fres.stage_best), "all_ranks_same_pose": same,
                       "collective": "ncclAllGather of the per-rank CC slices on device buffers, one per rotate() table (5 per fit)"
                                     if world > 1 else "none (one rank)",
                       "note": "whole vt_fit call incl. vol_com of the 256^3 target and the workspace set-up; max over ranks, best of 3"}

    # ---- the same cfg2 search as ONE call of the search API (vt_fit_search_grid): rotations x 27 translations, the
    # translated candidates of a rotation share that rotation's splat (the density of a translated structure is the same
    # voxels on a translated frame), ranks split the rotations, one ncclAllGather of (cc, index).  Reported beside the
    # headline, not as it: the headline scores an arbitrary pose list one splat per pose.
    grid_search = None
    if not args.no_extras:
        rx, ry, rz = np.arange(25) * (360.0 / 25), np.arange(20) * (360.0 / 20), np.arange(20) * (360.0 / 20)
        rots = np.stack(np.meshgrid(rx, ry, rz, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
        rots = np.ascontiguousarray(rots[np.random.default_rng(2004).permutation(len(rots))[:args.grid_rotations * world]])
        ctx.search_grid(com, rots[:64], 1, 1.0)             # warm-up
        barrier()
        t0 = time.perf_counter()
        gcc, gidx, gpose, _ = ctx.search_grid(com, rots, 1, 1.0)
        vt.sync()
        dt = time.perf_counter() - t0
        # a sample of the same candidates scored the per-pose way (own transform + splat each)
        os.environ["VOLTOOL_GRID_SHARED"] = "0"
        _, _, _, ref_all = ctx.search_grid(com, rots[:8], 1, 1.0, want_all=True) if world == 1 else (0, 0, 0, None)
        os.environ.pop("VOLTOOL_GRID_SHARED")
        sh_all = ctx.search_grid(com, rots[:8], 1, 1.0, want_all=True)[3] if world == 1 else None
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ncand = len(rots) * 27
        grid_search = {"rotations": len(rots), "translations": 27, "candidates": ncand, "seconds": float(tt[0].item()),
                       "poses_per_s": ncand / float(tt[0].item()), "ranks": world, "best_cc": float(gcc), "best_index": int(gidx),
                       "best_pose": [float(v) for v in gpose],
                       "max_rel_diff_vs_per_pose_path": (float(np.nanmax(np.abs(sh_all - ref_all) / np.maximum(np.abs(ref_all), 1e-3)))
                                                         if sh_all is not None else None),
                       "note": "vt_fit_search_grid with host rotation list in, winner out; one splat per rotation, 27 re-aimed cc "
                               "passes (batch_retarget); tol 1e-5 on CC against the per-pose path (tests/test_gpu_extras.py)"}

    # ---- cfg5's "fit with local rigid refinement on N GPUs": every rank holds the 250k-atom structure and the ~400^3
    # target; each refinement stage is sharded by the library and merged with one ncclAllGather of (cc, index)
    cfg5_sharded = None
    if world > 1 and not args.no_extras and args.cfg5_atoms > 0:
        from voltool_b200.synth import make_structure
        ball = 105.0 * (args.cfg5_atoms / 250000.0) ** (1 / 3)
        p5, r5, m5 = make_structure(args.cfg5_atoms, ball, seed=5001, center=(ball + 12.0,) * 3)
        rs5, _, q5 = vt.sim_policy(3.0)
        barrier()
        refine, ctx5 = cfg5_refine(vt, p5, r5, m5, q5, rs5, (2.0 * ball + 8.0) / args.cfg5_edge)
        ctx5.close()
        tt = torch.tensor([refine["seconds"]], device="cuda", dtype=torch.float64)
        sig = torch.tensor(refine["best_deg"] + [refine["best_cc"]], device="cuda", dtype=torch.float64)
        sig_all = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_gather(sig_all, sig)
        refine["seconds"] = float(tt[0].item())
        refine["poses_per_s"] = refine["poses"] / refine["seconds"]
        refine["ranks"] = world
        refine["all_ranks_same_pose"] = all(bool(torch.equal(x, sig_all[0])) for x in sig_all)
        refine["collective"] = "ncclAllGather of the per-rank (cc, index), one per stage"
        cfg5_sharded = refine
        del p5, r5, m5

    if rank != 0:
        if world > 1:
            vt.nccl_shutdown()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (splat): FP32/SFU bound, not HBM ----
    pairs = splat_pairs(radii, radscale, gsp)
    splat_ms, splat_n = prof["splat"]
    poses_done = P * K
    fp32_peak = props["sm_count"] * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
    achieved = FLOP_PER_PAIR * pairs * poses_done / (splat_ms * 1e-3) / 1e12 if splat_ms > 0 else None
    roofline = {"kernel": "splat_list_kernel", "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": (achieved / fp32_peak) if achieved else None,
                "traffic": splat_traffic(args, P * K / max(splat_n, 1))[0],
                "traffic_source": splat_traffic(args, P * K / max(splat_n, 1))[1],
                "peak_source": f"{props['sm_count']} SMs x 128 lanes x 2 x sm_max_mhz ({peaks['source']} MEASURED_PEAKS); measured on this "
                               "pool: 127.9 FMA/clk/SM with FFMA2 chains (profiles/r2_fp32_peak_microbench.txt); "
                               "no tensor cores: nothing here is a contraction",
                "binding_resource": "shared-memory pipe, then instruction issue (ncu of this kernel inside the fit step: l1tex 74 % "
                                    "busy, issue 72 %; profiles/r2l_splat_list_kernel_summary.txt); only ~1 in 7 (atom, voxel) "
                                    "evaluations of a warp lands inside the atom's cutoff box, which is what `frac` shows",
                "pairs_per_pose": pairs, "flop_per_pair": FLOP_PER_PAIR,
                "ms_per_launch": splat_ms / max(splat_n, 1), "launches": splat_n,
                "share_of_step": {k: v[0] / ms_total for k, v in prof.items()}}

    out = {"metric": "fit_poses_per_sec", "value": value, "unit": "poses/s", "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "cfg2 fit: 50k atoms into a 256^3 map, rotation x translation grid search",
                      "atoms": args.atoms, "map": [n, n, n], "resolution": RESOLUTION, "poses_per_step_per_gpu": P,
                      "pose_list": "25x20x20 Euler lattice x 3x3x3 shifts (270000), permuted, sliced per rank",
                      "l2": "per-step working set (target 64 MiB + per-pose grids/coordinates, several GB) exceeds the 126 MB L2",
                      "parallelism": f"pose-shard x{world}, one ncclAllGather of (cc, index) per step issued by the library"},
           "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": "poses/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": ms_e2e / K, "wall_s": wall_e2e},
           "gpu_launches": int(launches), "roofline": roofline,
           "best_of_last_step": {"cc": float(best[:, 0].max()), "planted_deg": PLANTED}}

    if fit_sharded is not None:
        out["fit_call"] = fit_sharded
    if cfg5_sharded is not None:
        out["cfg5_refine_sharded"] = cfg5_sharded
    if grid_search is not None:
        out["grid_search"] = grid_search
    if world == 1:
        if not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(pos, radii, com, tmap, poses, budget_s=args.cpu_budget)
            # parity of the benchmarked path: the poses the CPU reference arm just scored, scored again through the
            # timed entry point (device pose list -> vt_fit_eval_poses_device) and through the host-list entry point
            ref_cc = np.asarray(out["cpu_baseline"].pop("_ccs"), np.float64)
            m = len(ref_cc)
            d_p = torch.from_numpy(np.ascontiguousarray(poses[:m])).cuda()
            d_c = torch.empty(m, dtype=torch.float32, device="cuda")
            ctx.eval_poses_device(com, d_p.data_ptr(), m, d_c.data_ptr())
            vt.sync()
            got_dev = d_c.cpu().numpy().astype(np.float64)
            got_host = ctx.eval_poses(com, poses[:m]).astype(np.float64)
            rel = np.abs(got_dev - ref_cc) / np.maximum(np.abs(ref_cc), 1e-30)
            out["parity"] = {"n": int(m), "max_rel_err": float(rel.max()), "tol": 1e-5, "ok": bool(rel.max() <= 1e-5),
                             "device_list_equals_host_list_bits": bool(np.array_equal(got_dev, got_host)),
                             "against": "cpu_baseline's CC of the same poses (pose list entries incl. +-1 A shifts)",
                             "cc_gpu_first": float(got_dev[0]), "cc_cpu_first": float(ref_cc[0])}
        if not args.no_extras and not args.sharded_only:
            out["ops"] = extra_ops(vt, peaks, args)
            if args.cfg4_map > 0:
                out["cfg4"] = cfg4_legs(vt, peaks, args.cfg4_map)
                # the HBM-bound kernel of the path in the same shape as `roofline` (which describes the
                # FP32/shared-memory-bound splat): correlate of two resident maps, 8 B/voxel algorithmic
                c4 = out["cfg4"]["correlate"]
                nvox = float(args.cfg4_map) ** 3
                out["roofline_hbm"] = {"kernel": "cc_identity_kernel", "bound": "hbm", "achieved": c4["GB/s"],
                                       "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": c4["frac_hbm"],
                                       "traffic": 8.590082e9 + 3.345e6 if args.cfg4_map == 1024 else None,
                                       "traffic_source": "profiles/r1e_cc_identity_1024_summary.txt (dram read + write per launch); "
                                                         f"algorithmic {8.0 * nvox:.4g} B",
                                       "peak_source": f"{peaks['source']} MEASURED_PEAKS hbm_gbs", "ms_per_launch": c4["ms"]}
            if args.cfg5_atoms > 0:
                out["cfg5"] = cfg5_leg(vt, args.cfg5_atoms, args.cfg5_edge)
    print(json.dumps(out))
    if world > 1:
        vt.nccl_shutdown()
        dist.destroy_process_group()


def extra_ops(vt, peaks, args):
    """the other BASELINE metrics on one GPU: correlate / smooth as HBM GB/s, sim+cc as voxels/s"""
    import torch
    res = {}
    n = args.ops_map
    g = vt.Grid.make((0, 0, 0), (n, n, n), 1.0)
    rng = np.random.default_rng(4001)
    a = rng.random(g.n, dtype=np.float32)
    b = (0.7 * a + 0.3 * rng.random(g.n, dtype=np.float32)).astype(np.float32)
    A, B = vt.VolMap(g, a), vt.VolMap(g, b)

    def timeit(key, fn, reps=5):
        fn()
        vt.profile_reset(); vt.profile_enable(True)
        for _ in range(reps):
            fn()
        ms = sum(vt.profile_get(k)[0] for k in key.split("+"))
        vt.profile_enable(False)
        return ms / reps

    ms = timeit("cc", lambda: vt.cc_threaded(A, B))
    gbs = 8.0 * g.n / (ms * 1e-3) / 1e9
    res["correlate"] = {"map": [n, n, n], "ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_voxel": 8,
                        "cc": vt.cc_threaded(A, B)}
    res["correlate"]["path"] = "both maps on one exact lattice: streamed, no interpolation"
    # cfg4's second run: B offset by half a voxel in x forces real trilinear interpolation
    g_off = vt.Grid.make((0.5, 0, 0), (n, n, n), 1.0)
    B_off = vt.VolMap(g_off, b)
    ms = timeit("cc", lambda: vt.cc_threaded(A, B_off))
    gbs = 8.0 * g.n / (ms * 1e-3) / 1e9
    res["correlate_offset"] = {"map": [n, n, n], "ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"],
                               "bytes_per_voxel": 8, "cc": vt.cc_threaded(A, B_off),
                               "path": "B origin +0.5 voxel in x: real trilinear interpolation; regular sampling -> cc_lean_kernel "
                                       "(packed column pairs, tensor-map TMA ring)"}
    # both maps on ONE lattice whose spacing is not exactly representable (1.06 A): the float index arithmetic jitters,
    # some steps repeat or skip a source plane -> the general ring kernel
    g_j = vt.Grid.make((0, 0, 0), (n, n, n), 1.06)
    A_j, B_j = vt.VolMap(g_j, a), vt.VolMap(g_j, b)
    ms = timeit("cc", lambda: vt.cc_threaded(A_j, B_j))
    gbs = 8.0 * g.n / (ms * 1e-3) / 1e9
    res["correlate_jitter"] = {"map": [n, n, n], "ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_voxel": 8,
                               "cc": vt.cc_threaded(A_j, B_j), "path": "spacing 1.06 A on both maps: irregular index steps, general ring kernel"}
    del A_j, B_j
    ms = timeit("binary", lambda: vt.add(A, B))
    gbs = 12.0 * g.n / (ms * 1e-3) / 1e9
    res["add"] = {"ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_voxel": 12,
                  "path": "one exact lattice: streamed"}
    ms = timeit("binary", lambda: vt.add(A, B_off))
    gbs = 12.0 * g.n / (ms * 1e-3) / 1e9
    res["add_offset"] = {"ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_voxel": 12,
                         "path": "B origin +0.5 voxel in x: cc_lean_kernel (packed column pairs, tensor-map TMA ring)"}
    del B_off
    ms = timeit("unary", lambda: A.scalar_add(0.0))
    gbs = 8.0 * g.n / (ms * 1e-3) / 1e9
    res["sadd"] = {"ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_voxel": 8}
    ms = timeit("unary", lambda: A.clone().binmask(0.5))
    res["binmask"] = {"ms": ms, "GB/s": 8.0 * g.n / (ms * 1e-3) / 1e9, "bytes_per_voxel": 8}
    # `voltool sigma` = VolumetricData::sigma_scale (TclVoltool.C:2167): mean pass + sigma pass + scale pass = 16 B/voxel
    Sg = A.clone()
    ms = timeit("stats+unary", lambda: Sg.sigma_scale())
    gbs = 16.0 * g.n / (ms * 1e-3) / 1e9
    res["sigma"] = {"ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_voxel": 16,
                    "note": "sigma_scale: mean pass (4 B) + sigma pass (4 B) + scale pass (8 B)"}
    del Sg
    C = A.clone()
    ms = timeit("blur", lambda: C.gaussian_blur(2.0), reps=3)
    gbs = 8.0 * g.n / (ms * 1e-3) / 1e9
    res["smooth_sigma2"] = {"ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_voxel": 8,
                            "note": "13 taps: x+y fused per plane tile, then z: two passes move 16 B/voxel"}
    ms = timeit("blur", lambda: C.gaussian_blur(1.0), reps=3)
    gbs = 8.0 * g.n / (ms * 1e-3) / 1e9
    res["smooth_sigma1"] = {"ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_voxel": 8,
                            "note": "7 taps: ONE pass over HBM, tensor-map TMA staging (blur_tma_kernel)"}
    def resample_leg(op):
        vt.profile_reset(); vt.profile_enable(True)
        reps = 3
        for _ in range(reps):
            M = A.clone()
            getattr(M, op)()
            del M
        ms, cnt = vt.profile_get("resample")
        vt.profile_enable(False)
        return ms / reps
    getattr(A.clone(), "downsample")()
    ms = resample_leg("downsample")
    gbs = 4.5 * g.n / (ms * 1e-3) / 1e9
    res["downsample"] = {"ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_input_voxel": 4.5}
    if n <= 512:
        getattr(A.clone(), "supersample")()
        ms = resample_leg("supersample")
        gbs = 36.0 * g.n / (ms * 1e-3) / 1e9
        res["supersample"] = {"ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_input_voxel": 36,
                              "out": [2 * n - 1] * 3}
    del A, B, C
    res["sim_cc_cfg1"] = cfg1_leg(vt, args)
    return res


def cfg1_leg(vt, args):
    """BASELINE configs[0] as SURVEY 8d states it: 10 000 atoms (seed 1001, ball R = 30 A), -res 5 -spacing 1.0; the cc
    target is that sim + N(0, 0.05 max) noise (seed 1002) on the identical grid; thresholds -FLT_MAX and 0.1.
    One call = vt_sim_cc with host coordinates (atom sort + upload + splat + cc + readback).  Unit: voxels of the
    simulated map per second."""
    from voltool_b200.synth import make_structure, add_noise
    FLT_MAX = 3.4028234663852886e38
    pos, radii, _ = make_structure(10000, 30.0, seed=1001)
    radscale, _, quality = vt.sim_policy(5.0, 1.0)
    S = vt.sim(pos, radii, quality, radscale, 1.0)
    sim = S.data
    T = vt.VolMap(S.grid, add_noise(sim, 0.05, seed=1002))
    nvox = S.gridsize()
    out = {"voxels": nvox, "atoms": 10000, "target": "sim + N(0, 0.05 max), seed 1002, identical grid",
           "note": "vt_sim_cc through the C ABI with host coordinates (atom sort + upload + splat + cc + readback)"}
    reps = 5
    for name, thr in (("thr_minus_fltmax", -FLT_MAX), ("thr_0.1", 0.1)):
        vt.sim_cc(pos, radii, 5.0, T, threshold=thr, gspacing=1.0)
        vt.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            cc = vt.sim_cc(pos, radii, 5.0, T, threshold=thr, gspacing=1.0)
        vt.sync()
        dt = (time.perf_counter() - t0) / reps
        # the same call with the library's own events around its kernels: what the device spends
        vt.profile_reset(); vt.profile_enable(True)
        for _ in range(reps):
            vt.sim_cc(pos, radii, 5.0, T, threshold=thr, gspacing=1.0)
        vt.sync()
        kms = sum(vt.profile_get(k)[0] for k in ("transform", "setup", "splat", "cc")) / reps
        vt.profile_enable(False)
        out[name] = {"s_per_call": dt, "voxels_per_s": nvox / dt, "cc": cc, "kernel_ms_per_call": kms,
                     "kernel_voxels_per_s": nvox / (kms * 1e-3) if kms > 0 else None}
    if not args.no_cpu:
        # the reference's CPU path on the same input: restated splat (OpenMP) + the reference's cc_threaded
        try:
            lib, kind = load_reference()
            threads = cpu_threads()
            if kind == "reference":
                lib.set_numprocs(threads)
            from oracle.pyoracle import Grid as OGrid
            og = OGrid.make(tuple(S.grid.origin), (S.grid.xsize, S.grid.ysize, S.grid.zsize), 1.0)
            tdata = T.data
            cpu = {"cores": threads, "kind": kind}
            for name, thr in (("thr_minus_fltmax", -FLT_MAX), ("thr_0.1", 0.1)):
                t0 = time.perf_counter()
                sg, sd = lib.sim(pos, radii, quality, radscale, 1.0, nthreads=threads)
                ccr = lib.ref_cc(sg, sd, og, tdata, thr) if kind == "reference" else lib.cc(sg, sd, og, tdata, thr, nthreads=threads)
                dtc = time.perf_counter() - t0
                cpu[name] = {"s_per_call": dtc, "voxels_per_s": sg.n / dtc, "cc": ccr,
                             "rel_err_gpu_vs_cpu": abs(out[name]["cc"] - ccr) / max(abs(ccr), 1e-30)}
            out["cpu"] = cpu
        except Exception as e:  # the checker is optional for this leg
            out["cpu"] = {"unavailable": str(e)[:200]}
    return out


def cfg4_legs(vt, peaks, n):
    """BASELINE configs[3] at its own size: correlate + sigma / binmask / add / diff on two n^3 maps
    (default 1024^3 = 4 GiB each).  The maps are generated on the device (seeds 4001 / 4002) and handed
    to the library through vt_map_create_from_device; B = 0.7 A + 0.3 U' so the analytic CC is 0.919145."""
    import torch
    res = {"map": [n, n, n]}
    g = vt.Grid.make((0, 0, 0), (n, n, n), 1.0)
    gen = torch.Generator(device="cuda").manual_seed(4001)
    a = torch.rand(g.n, generator=gen, device="cuda", dtype=torch.float32)
    gen.manual_seed(4002)
    b = torch.rand(g.n, generator=gen, device="cuda", dtype=torch.float32).mul_(0.3).add_(a, alpha=0.7)
    torch.cuda.synchronize()
    A = vt.VolMap.from_device(g, a.data_ptr())
    B = vt.VolMap.from_device(g, b.data_ptr())
    g_off = vt.Grid.make((0.5, 0, 0), (n, n, n), 1.0)
    B_off = vt.VolMap.from_device(g_off, b.data_ptr())
    vt.sync()
    del a, b
    torch.cuda.empty_cache()

    def timeit(key, fn, reps=3):
        fn()
        vt.profile_reset(); vt.profile_enable(True)
        for _ in range(reps):
            fn()
        ms = sum(vt.profile_get(k)[0] for k in key.split("+"))
        vt.profile_enable(False)
        return ms / reps

    def leg(name, key, fn, bytes_per_voxel, **extra):
        ms = timeit(key, fn)
        gbs = bytes_per_voxel * g.n / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"], "bytes_per_voxel": bytes_per_voxel, **extra}

    leg("correlate", "cc", lambda: vt.cc_threaded(A, B), 8, cc=vt.cc_threaded(A, B), analytic_cc=0.919145,
        path="one exact lattice: streamed")
    leg("correlate_offset", "cc", lambda: vt.cc_threaded(A, B_off), 8, cc=vt.cc_threaded(A, B_off),
        path="B origin +0.5 voxel in x: real trilinear interpolation, cc_lean_kernel (packed column pairs, tensor-map TMA ring)")
    leg("add", "binary", lambda: vt.add(A, B), 12)
    leg("diff", "binary", lambda: vt.subtract(A, B), 12)
    leg("add_offset", "binary", lambda: vt.add(A, B_off), 12)
    del B_off
    leg("sigma", "stats+unary", lambda: A.sigma_scale(), 16,
        note="`voltool sigma` = sigma_scale (TclVoltool.C:2167): mean pass + sigma pass + scale pass; in place, repeated")
    leg("binmask", "unary", lambda: B.binmask(0.5), 8, note="in place (repeated on the masked map)")
    del A, B
    return res


def cfg5_leg(vt, natoms, edge):
    """BASELINE configs[4] on one GPU: ribosome-scale sim (natoms at 3 A into ~edge^3) and a local rigid
    refinement (+-12 deg in 4 deg steps, then +-3 deg in 1 deg steps about the winner) against a noisy map
    of a planted nearby pose.  Candidates shard across ranks exactly like the cfg2 search."""
    from voltool_b200.synth import make_structure, add_noise
    res = {}
    ball = 105.0 * (natoms / 250000.0) ** (1 / 3)
    pos, radii, mass = make_structure(natoms, ball, seed=5001, center=(ball + 12.0,) * 3)
    radscale, _, quality = vt.sim_policy(3.0)
    spacing = (2.0 * ball + 8.0) / edge
    vt.sim(pos, radii, quality, radscale, spacing)        # warm-up (allocations)
    vt.sync()
    t0 = time.perf_counter()
    S = vt.sim(pos, radii, quality, radscale, spacing)
    vt.sync()
    dt = time.perf_counter() - t0
    sg = S.grid
    pairs = splat_pairs(radii, radscale, spacing)
    res["sim"] = {"atoms": natoms, "map": [sg.xsize, sg.ysize, sg.zsize], "spacing": spacing, "seconds": dt,
                  "voxels_per_s": S.gridsize() / dt, "pairs": pairs, "pairs_per_s": pairs / dt,
                  "note": "vt_sim through the C ABI with host coordinates (sort + upload + splat + min/max)"}
    refine, ctx = cfg5_refine(vt, pos, radii, mass, quality, radscale, spacing)
    res["refine"] = refine
    ctx.close()
    return res


def cfg5_refine(vt, pos, radii, mass, quality, radscale, spacing):
    """the local rigid refinement of BASELINE configs[4]: two stages of 7^3 Euler triples (4 deg, then 1 deg about the
    winner) through vt_fit_search_grid.  With the library's communicator up (torchrun, N > 1) every rank calls this on
    the same inputs and the library shards each stage's candidates (rank r: r, r + N, ...) and merges the per-rank
    (cc, index) with one ncclAllGather per stage."""
    from voltool_b200.synth import add_noise
    com = vt.measure_center(pos, mass)
    planted = (8.0, -4.0, 5.0)
    T = vt.sim(vt.pose_apply(pos, com, planted), radii, quality, radscale, spacing)
    T.upload(add_noise(T.data, 0.02, seed=5002))
    ctx = vt.FitContext(pos, radii, T, 3.0)
    lat4 = np.arange(-12.0, 12.1, 4.0)
    lat1 = np.arange(-3.0, 3.1, 1.0)

    def two_stages():
        best, n_eval, best_cc = np.zeros(3), 0, float("nan")
        for lat in (lat4, lat1):
            rot = (best[None, :] + np.stack(np.meshgrid(lat, lat, lat, indexing="ij"), -1).reshape(-1, 3)).astype(np.float32)
            best_cc, idx, pose, _ = ctx.search_grid(com, rot, 0, 0.0)
            best = rot[idx].astype(np.float64)
            n_eval += len(rot)
        return best, n_eval, best_cc

    two_stages()                                           # warm-up: grows the pool by the pose workspaces
    vt.sync()
    t0 = time.perf_counter()
    best, n_eval, best_cc = two_stages()
    vt.sync()
    dt = time.perf_counter() - t0
    return ({"poses": n_eval, "seconds": dt, "poses_per_s": n_eval / dt, "best_deg": best.tolist(),
             "planted_deg": list(planted), "best_cc": float(best_cc), "batch": ctx.info()["batch"],
             "note": "two stages of 7^3 Euler triples (4 deg, then 1 deg about the winner) through vt_fit_search_grid; "
                     "the planted triple lies on that lattice"}, ctx)


# ================================================================================================ reference arm
def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_reference():
    """oracle/_ref (the reference's own object code) when present, else the C port"""
    from oracle import pyoracle
    if pyoracle.ref_available():
        return pyoracle.Ref(), "reference"
    return pyoracle.Oracle(), "port"


def cpu_eval_poses(lib, kind, pos, radii, com, g, t, tvol, poses, threads):
    """calc_cc per pose on the host cores: restated splat (OpenMP) + cc_threaded"""
    if kind == "reference":
        lib.set_numprocs(threads)
        lib.install_cc_backend(tvol)
    out = []
    for q in poses:
        p = pose_host(lib, pos, com, q)
        out.append(lib.calc_cc(p, radii, g, t, RESOLUTION, nthreads=threads))
    return out


def pose_host(lib, pos, com, q):
    """one candidate of the pose list on the CPU: the oracle's restatement of rotate()'s transform
    (TclVoltool.C:157-176) for float degrees, then the shift"""
    return lib.pose_general(pos, com, q[:3], q[3:6]).reshape(-1, 3)


def cpu_baseline(pos, radii, com, tmap, poses, budget_s=20.0):
    """rank 0, N=1: the reference CPU path timed on a bounded sample of the same pose list"""
    from oracle.pyoracle import Grid as OGrid
    lib, kind = load_reference()
    g = OGrid()
    tg = tmap.grid
    for k in range(3):
        g.origin[k], g.xaxis[k], g.yaxis[k], g.zaxis[k] = tg.origin[k], tg.xaxis[k], tg.yaxis[k], tg.zaxis[k]
    g.xsize, g.ysize, g.zsize = tg.xsize, tg.ysize, tg.zsize
    t = tmap.data
    threads = cpu_threads()
    tvol = lib.vol(g, t) if kind == "reference" else None
    done, t0 = 0, time.perf_counter()
    ccs = []
    while done < 64:
        ccs += cpu_eval_poses(lib, kind, pos, radii, com, g, t, tvol, poses[done:done + 2], threads)
        done += 2
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    if kind == "reference":
        lib.install_cc_backend(None)
    return {"value": done / dt, "unit": "poses/s", "cores": threads, "kind": kind,
            "sample": f"{done} poses of the same list, {dt:.1f} s; splat = documented restatement (QuickSurf not in the "
                      f"snapshot, OpenMP z-slabs), cc = {'the reference cc_threaded object code' if kind == 'reference' else 'C port of cc_threaded'}",
            "cc_first": float(ccs[0]), "_ccs": [float(c) for c in ccs]}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle.pyoracle import Grid as OGrid  # noqa: F401
    from voltool_b200.synth import make_structure, add_noise
    lib, kind = load_reference()
    threads = cpu_threads()
    n = args.map
    pos, radii, mass = make_structure(args.atoms, 52.0 * (args.atoms / 50000.0) ** (1 / 3), seed=2001,
                                      center=(0.5 * (n - 1),) * 3)
    com0 = lib.measure_center(pos, mass)
    planted = pose_host(lib, pos, com0, np.array(PLANTED + (0.0, 0.0, 0.0), np.float32))
    radscale, gsp, quality = lib.sim_policy(RESOLUTION)
    g, d = lib.sim(planted, radii, quality, radscale, 1.0, nthreads=threads)
    padm = [(n - s) // 2 for s in (g.xsize, g.ysize, g.zsize)]
    padp = [n - s - m for s, m in zip((g.xsize, g.ysize, g.zsize), padm)]
    g, d = lib.pad(g, d, (padm[0], padp[0], padm[1], padp[1], padm[2], padp[2]))
    d = add_noise(d, 0.02, seed=2002)
    dencom = lib.vol_com(g, d)
    pos = (pos + (dencom - com0)[None, :]).astype(np.float32)
    com = lib.measure_center(pos, mass)
    tvol = lib.vol(g, d) if kind == "reference" else None
    poses = pose_list()
    S = args.ref_poses
    W, K = args.warmup, args.steps

    def step(s):
        sl = poses[(s * S) % len(poses):][:S]
        return cpu_eval_poses(lib, kind, pos, radii, com, g, d, tvol, sl, threads)

    for s in range(W):
        step(s)
    t0 = time.perf_counter()
    for s in range(W, W + K):
        cc = step(s)
    dt = time.perf_counter() - t0
    value = S * K / dt
    out = {"impl": "reference", "metric": "fit_poses_per_sec", "value": value, "unit": "poses/s", "n_gpus": args.gpus,
           "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "cfg2 fit: 50k atoms into a 256^3 map, rotation x translation grid search",
                      "atoms": args.atoms, "map": [n, n, n], "resolution": RESOLUTION, "poses_per_step": S,
                      "note": "bounded sample of the same pose list; CPU only, rank 0"},
           "cpu_baseline": {"value": value, "unit": "poses/s", "cores": threads, "kind": kind,
                            "sample": f"{S} poses per step; splat = documented restatement (OpenMP), "
                                      f"cc = {'reference cc_threaded object code (oracle/_ref)' if kind == 'reference' else 'C port'}"},
           "e2e": {"value": value, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "last_cc": float(cc[-1])}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--atoms", type=int, default=50000)
    ap.add_argument("--map", type=int, default=256)
    ap.add_argument("--poses", type=int, default=2048, help="poses per step per GPU")
    ap.add_argument("--ref-poses", type=int, default=4, help="poses per step of the CPU reference arm")
    ap.add_argument("--ops-map", type=int, default=512, help="edge of the maps used by the extra ops legs")
    ap.add_argument("--cfg4-map", type=int, default=1024, help="edge of the cfg4 maps (0: skip the leg)")
    ap.add_argument("--cfg5-atoms", type=int, default=250000, help="atoms of the cfg5 leg (0: skip)")
    ap.add_argument("--cfg5-edge", type=int, default=400, help="target edge of the cfg5 sim grid")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--grid-rotations", type=int, default=1024, help="rotations per GPU of the grid_search leg (x 27 translations)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--sharded-only", action="store_true",
                    help="keep the all-rank legs (fit_call, cfg5_refine_sharded) but skip rank 0's single-GPU legs and the CPU baseline")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3   # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
