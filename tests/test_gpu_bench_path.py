"""GPU parity of the BENCHMARKED path against the CPU oracle, at BASELINE.json's own configurations:

  * the device pose list (vt_fit_eval_poses_device, what bench.py times) == the host pose list, bit for bit,
    on non-lattice angles and shifted poses; shifted poses vs the oracle's calc_cc;
  * cfg1 exactly as BASELINE.md / SURVEY 8d state it (10k atoms, seed 1001, 1 A spacing, target = sim + N(0, 0.05 max)
    seed 1002, thresholds -FLT_MAX and 0.1) through vt_sim_cc incl. synth_out;
  * 8 poses of bench.py's cfg2 list (50k atoms, 256^3) vs the oracle;
  * cfg5's 250k-atom sim: identical support and 1e-5 on every voxel;
  * the host-buffer entry points vt_cc_host / vt_binary_host and vt_map_create_from_device.

Tolerances are the north star's: 1e-5 relative for sim voxels and CC, bit-exact for everything integer or elementwise.
"""
import os

import numpy as np
import pytest

import voltool_b200 as vt
from oracle.pyoracle import Grid as OGrid
from tests.synth import make_structure, make_target, add_noise

pytestmark = pytest.mark.gpu
FLT_MAX = 3.4028234663852886e38
TOL = 1e-5
NT = max(1, min(16, len(os.sched_getaffinity(0))))


def G(g):
    return vt.Grid.from_fields(g)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module", autouse=True)
def _init():
    vt.init(0)
    yield


def random_poses(n, seed, shift=1.5):
    r = np.random.default_rng(seed)
    rot = (r.random((n, 3)) * 720.0 - 360.0).astype(np.float32)      # non-lattice, both signs, beyond one turn
    sh = ((r.random((n, 3)) * 2 - 1) * shift).astype(np.float32)
    p = np.concatenate([rot, sh], 1).astype(np.float32)
    p[0] = 0.0                                                        # identity
    p[1, :3] = (90.0, 180.0, 270.0); p[1, 3:] = 0.0                   # cos/sin at their zeros
    p[2, :3] = (1e-3, -1e-3, 359.999)
    return np.ascontiguousarray(p)


# ------------------------------------------------------------------------------------------ device list == host list
@pytest.mark.parametrize("force_host_prep", [False, True])
def test_device_pose_list_equals_host_list(oracle, monkeypatch, force_host_prep):
    import torch
    pos, radii, mass = make_structure(1500, 14.0, seed=61)
    g, t = make_target(oracle, pos, radii, 5.0, 1.0, rot=(24, 48, 0), noise=0.02, seed=62, mass=mass)
    target = vt.VolMap(G(g), t)
    com = vt.measure_center(pos, mass)
    ctx = vt.FitContext(pos, radii, target, 5.0)
    poses = random_poses(700, 63)                                     # three batches of 256
    want = ctx.eval_poses(com, poses)
    if force_host_prep:
        monkeypatch.setenv("VOLTOOL_PREP_HOST", "1")
    d_p = torch.from_numpy(poses).cuda()
    d_c = torch.full((len(poses),), -7.0, dtype=torch.float32, device="cuda")
    ctx.eval_poses_device(com, d_p.data_ptr(), len(poses), d_c.data_ptr())
    vt.sync()
    got = d_c.cpu().numpy()
    assert np.array_equal(bits(got), bits(want))
    if force_host_prep:
        assert ctx.info()["prep_on_host"]


def test_device_cos_sin_match_host_libm():
    """prep_from_pose_kernel against euler_cs on a dense sweep of angles: every pose the device kernel does not
    flag as ambiguous must have the host's bits (checked through the transform of one atom, which exposes c and s)"""
    import torch
    pos = np.array([[1.0, 2.0, 3.0], [-2.0, 0.5, 1.0]], np.float32)
    radii = np.array([1.5, 1.2], np.float32)
    g = vt.Grid.make((-12, -12, -12), (24, 24, 24), 1.0)
    target = vt.VolMap(g, np.random.default_rng(5).random(g.n, dtype=np.float32))
    com = np.zeros(3, np.float32)
    ctx = vt.FitContext(pos, radii, target, 5.0)
    n = 4096
    r = np.random.default_rng(64)
    poses = np.zeros((n, 6), np.float32)
    poses[:, :3] = (r.random((n, 3)) * 360.0).astype(np.float32)
    poses[: n // 2, :3] = np.round(poses[: n // 2, :3])               # whole degrees like the fit lattices
    want = ctx.eval_poses(com, poses)
    d_p = torch.from_numpy(poses).cuda()
    d_c = torch.empty(n, dtype=torch.float32, device="cuda")
    ctx.eval_poses_device(com, d_p.data_ptr(), n, d_c.data_ptr())
    vt.sync()
    assert np.array_equal(bits(d_c.cpu().numpy()), bits(want))


# ------------------------------------------------------------------------------------------ shifted poses vs oracle
def test_shifted_poses_match_oracle(oracle):
    pos, radii, mass = make_structure(1200, 13.0, seed=71)
    g, t = make_target(oracle, pos, radii, 5.0, 1.0, rot=(24, 48, 0), noise=0.02, seed=72, mass=mass, pad=5)
    target = vt.VolMap(G(g), t)
    com = vt.measure_center(pos, mass)
    ctx = vt.FitContext(pos, radii, target, 5.0)
    poses = random_poses(24, 73, shift=2.0)
    poses[3] = (24.0, 48.0, 0.0, 1.0, -1.0, 1.0)                      # a bench-list style entry: lattice angles, +-1 A
    got = ctx.eval_poses(com, poses)
    for q, c in zip(poses, got):
        p = oracle.pose_general(pos, com, q[:3], q[3:])
        assert np.array_equal(bits(vt.pose_apply(pos, com, q[:3], q[3:])), bits(p.reshape(-1, 3)))
        w = float(np.float32(oracle.calc_cc(p, radii, g, t, 5.0, nthreads=NT)))
        assert abs(float(c) - w) <= TOL * abs(w), (q, c, w)


# ------------------------------------------------------------------------------------------ cfg1 as BASELINE states it
def test_cfg1_sim_cc_as_baseline_states_it(oracle):
    pos, radii, _ = make_structure(10000, 30.0, seed=1001)            # ball R = 30 A centred at (50, 50, 50)
    radscale, gsp, quality = oracle.sim_policy(5.0, 1.0)              # -res 5 -spacing 1.0
    assert (radscale, gsp, quality) == (1.0, 1.0, 3)
    g, sim = oracle.sim(pos, radii, quality, radscale, gsp, nthreads=NT)
    tgt = add_noise(sim, 0.05, seed=1002)                             # same sim + N(0, 0.05 max), identical grid
    target = vt.VolMap(G(g), tgt)
    for thr in (-FLT_MAX, 0.1):
        want, nwant, _ = oracle.cc(g, sim, g, tgt, thr, nthreads=1, full=True)
        cc, synth = vt.sim_cc(pos, radii, 5.0, target, threshold=thr, gspacing=1.0, want_synth=True)
        assert abs(cc - want) <= TOL * abs(want), (thr, cc, want)
        assert synth.grid.astuple() == g.astuple()
        got = synth.data
        assert np.array_equal(got == 0, sim == 0)                     # identical support
        rel = np.abs(got - sim) / np.maximum(np.abs(sim), 1e-30)
        assert rel.max() < TOL
        # the synthetic map that comes back is the map that was scored
        cc2, n2, _ = vt.cc_threaded(synth, target, thr, full=True)
        assert cc2 == cc
        # the number of voxels past the threshold differs from the oracle's only where a voxel sits within 1e-5 of it
        assert abs(n2 - nwant) <= (0 if thr < 0 else 40)
        mn, mx = synth.datarange()
        assert mn == got.min() and mx == got.max()
    # the selection form of the same call: atoms listed among unselected ones
    r = np.random.default_rng(3)
    on = (r.random(15000) < 0.5).astype(np.int32)
    on[np.flatnonzero(on)[10000:]] = 0
    k = int(on.sum())
    full_pos = r.random((15000, 3)).astype(np.float32) * 100.0
    full_rad = np.full(15000, 9.0, np.float32)
    full_pos[on != 0] = pos[:k]
    full_rad[on != 0] = radii[:k]
    assert vt.sim_cc_sel(full_pos, full_rad, on, 5.0, target, -FLT_MAX, 1.0) == vt.sim_cc(pos[:k], radii[:k], 5.0, target,
                                                                                           gspacing=1.0)


# ------------------------------------------------------------------------------------------ cfg2: poses of the bench list
def test_cfg2_bench_poses_match_oracle(oracle):
    import bench
    pos, radii, mass, com, tmap = bench.build_cfg2(vt, 50000, 256)
    tg, t = tmap.grid, tmap.data
    og = OGrid.make(tuple(tg.origin), (tg.xsize, tg.ysize, tg.zsize), 1.0)
    assert og.astuple() == tg.astuple()
    ctx = vt.FitContext(pos, radii, tmap, bench.RESOLUTION)
    poses = bench.pose_list()[:8]
    assert np.any(poses[:, 3:] != 0)                                  # the list's entries carry +-1 A shifts
    got = ctx.eval_poses(com, poses)
    for q, c in zip(poses, got):
        p = oracle.pose_general(pos, com, q[:3], q[3:])
        w = float(np.float32(oracle.calc_cc(p, radii, og, t, bench.RESOLUTION, nthreads=NT)))
        assert abs(float(c) - w) <= TOL * abs(w), (q, c, w)
    # ... and the planted pose scores far above them
    best = ctx.eval_poses(com, vt.FitContext.make_poses(np.array([bench.PLANTED], np.float32)))
    assert best[0] > got.max() + 0.2


# ------------------------------------------------------------------------------------------ cfg5: 250k-atom sim
def test_cfg5_sim_250k_matches_oracle(oracle):
    natoms, edge = 250000, 400
    ball = 105.0
    pos, radii, _ = make_structure(natoms, ball, seed=5001, center=(ball + 12.0,) * 3)
    radscale, _, quality = oracle.sim_policy(3.0)
    spacing = (2.0 * ball + 8.0) / edge
    g, want = oracle.sim(pos, radii, quality, radscale, spacing, nthreads=NT)
    m = vt.sim(pos, radii, quality, radscale, spacing)
    assert m.grid.astuple() == g.astuple()
    assert min(g.xsize, g.ysize, g.zsize) >= 390
    got = m.data
    assert np.array_equal(got == 0, want == 0)
    nz = want != 0
    assert (np.abs(got[nz] - want[nz]) / np.abs(want[nz])).max() < TOL


# ------------------------------------------------------------------------------------------ host-buffer entry points
def test_cc_host_and_binary_host_and_from_device(oracle):
    import torch
    A = OGrid.make((0.0, 0.0, 0.0), (40, 36, 30), 1.0)
    B = OGrid.make((0.5, -1.25, 2.0), (37, 41, 33), (1.1, 0.9, 1.3))
    r = np.random.default_rng(81)
    a, b = r.random(A.n, dtype=np.float32), r.random(B.n, dtype=np.float32)
    for thr in (-FLT_MAX, 0.3):
        want = oracle.cc(A, a, B, b, thr)
        got = vt.cc_host(G(A), a, G(B), b, thr)                       # the body INTEGRATION.md gives cc_threaded
        assert abs(got - want) <= 1e-9
        assert got == vt.cc_threaded(vt.VolMap(G(A), a), vt.VolMap(G(B), b), thr)
    for op in range(4):
        for interp in (True, False):
            for union in (False, True):
                go, want = oracle.binary(op, A, a, B, b, interp, union)
                gg, got = vt.binary_host(op, G(A), a, G(B), b, interp, union)
                assert gg.astuple() == go.astuple()
                assert np.array_equal(bits(got), bits(want)), (op, interp, union)
    # a map adopted from a device pointer is the same map: data, min/max cache, CC
    d = torch.from_numpy(a).cuda()
    M = vt.VolMap.from_device(G(A), d.data_ptr())
    vt.sync()
    del d
    assert np.array_equal(bits(M.data), bits(a))
    c = M.cache
    assert c.minmax_valid == 1 and c.cmin == a.min() and c.cmax == a.max()
    assert vt.cc_threaded(M, vt.VolMap(G(B), b)) == vt.cc_threaded(vt.VolMap(G(A), a), vt.VolMap(G(B), b))


# ------------------------------------------------------------------------------------------ selections
def test_fit_sel_equals_fit_on_compacted(oracle):
    pos, radii, mass = make_structure(400, 9.0, seed=41)
    g, t = make_target(oracle, pos, radii, 6.0, 1.2, rot=(48, 24, 312), noise=0.01, seed=42, mass=mass, pad=6)
    target = vt.VolMap(G(g), t)
    r = np.random.default_rng(43)
    n_all = 700
    on = np.zeros(n_all, np.int32)
    on[np.sort(r.choice(np.arange(20, 680), 400, replace=False))] = 1
    full_pos = (r.random((n_all, 3)) * 80).astype(np.float32)
    full_rad = np.full(n_all, 3.0, np.float32)
    full_mass = np.full(n_all, 99.0, np.float32)
    full_pos[on != 0], full_rad[on != 0], full_mass[on != 0] = pos, radii, mass
    p_sel, r_sel = vt.fit_sel(full_pos, full_rad, full_mass, on, target, 6.0)
    p_ref, r_ref = vt.fit(pos, radii, mass, target, 6.0)
    assert np.array_equal(bits(p_sel[on != 0]), bits(p_ref))
    assert np.array_equal(bits(p_sel[on == 0]), bits(full_pos[on == 0]))   # unselected atoms untouched
    assert r_sel.best_cc == r_ref.best_cc and list(r_sel.stage_best) == list(r_ref.stage_best)
    S = vt.sim_sel(full_pos, full_rad, on, 3, 1.0, 1.0)
    assert np.array_equal(bits(S.data), bits(vt.sim(pos, radii, 3, 1.0, 1.0).data))
