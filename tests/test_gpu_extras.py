"""GPU parity of the SURVEY 8f rows against their oracle twins:

  f1  `voltool cc -calcdiffmap / -calcspatialccmap` (TclMDFF.C:180-184, 555-594): vt_cc_maps / vt_sim_cc_maps
  f4  the map-rotating search density_rotate / reset_density (TclVoltool.C:202-311) and the rotation x translation
      grid search (vt_fit_search_grid)

Both maps and both searches are restatements (the reference's own code for them is absent or dead, see the headers
of vt_ccmaps.cu / vt_density.cu and oracle/voltool_oracle.c); the tests pin the CUDA path to the CPU twin.
"""
import os

import numpy as np
import pytest

import voltool_b200 as vt
from oracle.pyoracle import Grid as OGrid
from tests.synth import make_structure, make_target

pytestmark = pytest.mark.gpu
FLT_MAX = 3.4028234663852886e38
NT = max(1, min(16, len(os.sched_getaffinity(0))))


def G(g):
    return vt.Grid.from_fields(g)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_grid(g1, g2):
    t1, t2 = g1.astuple(), g2.astuple()
    return t1[4] == t2[4] and all(np.array_equal(np.array(a), np.array(b)) for a, b in zip(t1[:4], t2[:4]))


@pytest.fixture(scope="module", autouse=True)
def _init():
    vt.init(0)
    yield


def rng_map(seed, n):
    return np.random.default_rng(seed).random(n, dtype=np.float32)


# ------------------------------------------------------------------------------------------ f1: diff + spatial CC maps
PAIRS = [
    (OGrid.make((0, 0, 0), (20, 18, 16), 1.0), OGrid.make((0, 0, 0), (20, 18, 16), 1.0)),                 # one lattice
    (OGrid.make((0.5, -1.25, 2.0), (17, 21, 13), (1.1, 0.9, 1.3)), OGrid.make((0, 0, 0), (20, 18, 16), 1.0)),
    (OGrid.make((-3.3, 1.7, 0.4), (25, 12, 19), 0.77), OGrid.make((1.0e-3, 2.5, -2.0), (9, 31, 14), (2.0, 0.5, 1.0))),
    (OGrid.make((0, 0, 0), (8, 8, 8), 1.0), OGrid.make((0.25, 0.25, 0.25), (33, 9, 7), 0.5)),             # one tile / ragged
    (OGrid.make((0.3, 0.1, -0.2), (15, 14, 13), axes=((14.0, 1.0, 0.0), (0.5, 13.0, 0.7), (0.0, 0.3, 12.0))),
     OGrid.make((0, 0, 0), (20, 18, 16), 1.0)),                                                             # skewed cell
]


@pytest.mark.parametrize("pair", range(len(PAIRS)))
@pytest.mark.parametrize("thr", [-FLT_MAX, 0.4])
def test_cc_maps_equal_oracle(oracle, pair, thr):
    gA, gB = PAIRS[pair]
    a, b = rng_map(11 + pair, gA.n), rng_map(29 + pair, gB.n)
    if pair == 2:
        a[5] = np.nan; b[17] = np.nan                      # NaN voxels drop out of the sums, propagate into diff
    cc_o, dg_o, d_o, sg_o, s_o = oracle.cc_maps(gA, a, gB, b, thr)
    A, B = vt.VolMap(G(gA), a), vt.VolMap(G(gB), b)
    cc, diff, spat = vt.cc_maps(A, B, thr)
    assert same_grid(diff.grid, G(dg_o)) and same_grid(spat.grid, G(sg_o))
    dd, ss = diff.data.reshape(-1), spat.data.reshape(-1)
    assert np.array_equal(np.isnan(dd), np.isnan(d_o))
    assert np.array_equal(bits(dd)[~np.isnan(d_o)], bits(d_o)[~np.isnan(d_o)])      # a - b: bit-exact
    assert np.array_equal(bits(ss), bits(s_o))                                      # tile sums in voxel order: bit-exact
    assert (np.isnan(cc) and np.isnan(cc_o)) or cc == cc_o
    # and the global value is cc_threaded's, up to the order of the double additions
    cc_ref = oracle.cc(gA, a, gB, b, thr)
    assert (np.isnan(cc) and np.isnan(cc_ref)) or abs(cc - cc_ref) <= 1e-9
    # the outputs are ordinary maps: the constructor's min/max is cached
    c = diff.cache
    finite = dd[~np.isnan(dd)]
    if finite.size == dd.size and dd.size:
        assert c.minmax_valid and c.cmin == dd.min() and c.cmax == dd.max()
    # only one of the two requested
    cc2, d2, s2 = vt.cc_maps(A, B, thr, want_diff=False)
    assert d2 is None and np.array_equal(bits(s2.data), bits(ss))


def test_sim_cc_maps_cfg1_like(oracle):
    """mdff_cc with all three map outputs on a cfg1-shaped case (2k atoms, 1 A spacing, noisy target)"""
    pos, radii, mass = make_structure(2000, 16.0, seed=1001)
    radscale, gsp, quality = oracle.sim_policy(5.0, 1.0)
    sg, sd = oracle.sim(pos, radii, quality, radscale, gsp, nthreads=NT)
    rng = np.random.default_rng(1002)
    t = (sd + rng.normal(0.0, 0.05 * sd.max(), sd.shape)).astype(np.float32)
    target = vt.VolMap(G(sg), t)
    for thr in (-FLT_MAX, 0.1):
        cc, synth, diff, spat = vt.sim_cc_maps(pos, radii, 5.0, target, thr, 1.0)
        s_gpu = synth.data.reshape(-1)
        assert np.array_equal(s_gpu != 0, sd != 0)
        assert np.max(np.abs(s_gpu - sd) / np.maximum(np.abs(sd), 1e-30)) < 1e-5
        # the maps follow from the GPU's own synth map exactly as the oracle derives them from it
        cc_o, dg_o, d_o, sg_o, s_o = oracle.cc_maps(sg, s_gpu, sg, t, thr)
        assert np.array_equal(bits(diff.data.reshape(-1)), bits(d_o))
        assert np.array_equal(bits(spat.data.reshape(-1)), bits(s_o))
        assert cc == cc_o
        assert abs(cc - vt.sim_cc(pos, radii, 5.0, target, thr, 1.0)) <= 1e-9
        assert abs(cc - oracle.cc(sg, sd, sg, t, thr)) <= 1e-5 * abs(cc)


# ------------------------------------------------------------------------------------------ f4: map-rotating search
def synth_and_target(oracle, n=600, rot=(0, 30, 60)):
    pos, radii, mass = make_structure(n, 10.0, seed=71)
    radscale, gsp, quality = oracle.sim_policy(6.0)
    sg, sd = oracle.sim(pos, radii, quality, radscale, gsp, nthreads=NT)
    gT, t = make_target(oracle, pos, radii, 6.0, 1.5, rot=rot, noise=0.01, seed=72, mass=mass)
    return pos, radii, mass, sg, sd, gT, t


@pytest.mark.parametrize("stride,max_rot", [(30, 3), (-15, 2), (90, 4)])
def test_density_rotate_equals_oracle(oracle, stride, max_rot):
    pos, radii, mass, sg, sd, gT, t = synth_and_target(oracle)
    com = oracle.vol_com(sg, sd)
    cc_o, best_o, tab_o = oracle.density_rotate(gT, t, sg, sd, stride, max_rot, com, -1.0, nthreads=NT)
    T, S = vt.VolMap(G(gT), t), vt.VolMap(G(sg), sd)
    g_before = S.grid.astuple()
    cc, best, tab = vt.density_rotate(T, S, stride, max_rot, com, -1.0)
    assert S.grid.astuple() == g_before                                    # frame restored (TclVoltool.C:262-273)
    ok = ~np.isnan(tab_o)
    assert np.array_equal(np.isnan(tab), np.isnan(tab_o))
    assert np.max(np.abs(tab[ok] - tab_o[ok])) <= 1e-6                     # float(cc): double sums differ in order only
    assert best == best_o and abs(cc - cc_o) <= 1e-6
    assert best is not None
    # a running best that nothing beats is left alone
    cc2, best2, _ = vt.density_rotate(T, S, stride, max_rot, com, 2.0)
    assert cc2 == 2.0 and best2 is None


def test_density_reset_and_pose_equal_oracle(oracle):
    pos, radii, mass, sg, sd, gT, t = synth_and_target(oracle)
    synthcom = oracle.vol_com(sg, sd)
    com = vt.measure_center(pos, mass)
    g_o = oracle.density_reset(sg, sd, 15, (1, 2, 3), synthcom, com)
    S = vt.VolMap(G(sg), sd)
    vt.density_reset(S, 15, (1, 2, 3), synthcom, com)
    assert same_grid(S.grid, G(g_o))                                       # metadata + exact vol_com: bit-equal frame
    # the atom half of reset_density (:302-311) is one candidate transform
    p_o = oracle.pose_transform(pos, com, 15, 1, 2, 3)
    p = vt.pose_apply(pos, com, (15.0, 30.0, 45.0))
    assert np.array_equal(bits(p), bits(p_o))


# ------------------------------------------------------------------------------------------ f4: rotation x translation
def test_search_grid_equals_oracle(oracle):
    pos, radii, mass = make_structure(500, 10.0, seed=81)
    gT, t = make_target(oracle, pos, radii, 5.0, 1.0, rot=(24, 0, 48), noise=0.02, seed=82, mass=mass)
    # plant a translation too: the target frame is shifted by (+1, 0, -1) A
    gT.origin[0] += 1.0; gT.origin[2] -= 1.0
    target = vt.VolMap(G(gT), t)
    com = vt.measure_center(pos, mass)
    rots = np.array([[0, 0, 0], [24, 0, 48], [24, 24, 48], [48, 0, 24]], np.float32)
    tsteps, tstep = 1, np.float32(1.0)
    ctx = vt.FitContext(pos, radii, target, 5.0)
    bcc, bidx, bpose, allcc = ctx.search_grid(com, rots, tsteps, float(tstep), want_all=True)
    nt = 2 * tsteps + 1
    exp = np.empty(len(rots) * nt ** 3, np.float32)
    for ir, r in enumerate(rots):
        for ix in range(nt):
            for iy in range(nt):
                for iz in range(nt):
                    sh = np.array([np.float32(ix - tsteps) * tstep, np.float32(iy - tsteps) * tstep,
                                   np.float32(iz - tsteps) * tstep], np.float32)
                    p = oracle.pose_general(pos, com, r, sh)
                    exp[((ir * nt + ix) * nt + iy) * nt + iz] = np.float32(oracle.calc_cc(p, radii, gT, t, 5.0, nthreads=NT))
    assert np.max(np.abs(allcc - exp) / np.abs(exp)) < 1e-5
    best = -np.inf; bi = -1
    for i, v in enumerate(exp):                                            # `if (cc > best)`: first maximum
        if v > best:
            best, bi = v, i
    assert bidx == bi and abs(bcc - best) <= 1e-5 * abs(best)
    ir, it = divmod(bi, nt ** 3)
    assert np.array_equal(bpose[:3], rots[ir])
    assert tuple(bpose[3:]) == tuple(np.float32(v - tsteps) * tstep for v in (it // (nt * nt), (it // nt) % nt, it % nt))
    # the planted pose wins: rotation (24, 0, 48), shift (+1, 0, -1)
    assert ir == 1 and tuple(bpose[3:]) == (1.0, 0.0, -1.0)
    ctx.close()


def test_search_grid_shared_splat_equals_per_pose_path(oracle, monkeypatch):
    """translated candidates of one rotation share that rotation's splat (batch_retarget): every score must agree with
    the path that transforms and splats each candidate separately (1e-5 on CC; measured ~1e-7), same winner"""
    pos, radii, mass = make_structure(3000, 18.0, seed=83)
    gT, t = make_target(oracle, pos, radii, 5.0, 1.0, rot=(48, 24, 0), noise=0.02, seed=84, mass=mass)
    gT.origin[1] += 1.5
    target = vt.VolMap(G(gT), t)
    com = vt.measure_center(pos, mass)
    r = np.random.default_rng(85)
    rots = np.concatenate([np.array([[48, 24, 0], [0, 0, 0]], np.float32), (r.random((300, 3)) * 360).astype(np.float32)])
    ctx = vt.FitContext(pos, radii, target, 5.0)
    monkeypatch.setenv("VOLTOOL_GRID_SHARED", "0")
    c0, i0, p0, all0 = ctx.search_grid(com, rots, 1, 1.5, want_all=True)
    monkeypatch.delenv("VOLTOOL_GRID_SHARED")
    c1, i1, p1, all1 = ctx.search_grid(com, rots, 1, 1.5, want_all=True)
    ok = ~np.isnan(all0)
    assert np.array_equal(np.isnan(all1), np.isnan(all0)) and ok.sum() > 8000
    rel = np.abs(all1[ok] - all0[ok]) / np.maximum(np.abs(all0[ok]), 1e-3)
    assert rel.max() < 1e-5, rel.max()
    assert i1 == i0 and np.array_equal(p1, p0) and abs(c1 - c0) <= 1e-5 * abs(c0)
    assert i1 // 27 == 0 and tuple(p1[3:]) == (0.0, 1.5, 0.0)              # the planted rotation and shift
    ctx.close()


def test_new_entry_points_edge_cases(oracle):
    """error behaviour and degenerate arguments of the round-2 entry points"""
    gA, gB = OGrid.make((0, 0, 0), (8, 8, 8), 1.0), OGrid.make((100, 100, 100), (8, 8, 8), 1.0)
    A, B = vt.VolMap(G(gA), rng_map(1, gA.n)), vt.VolMap(G(gB), rng_map(2, gB.n))
    with pytest.raises(vt.VoltoolGpuError):
        vt.cc_maps(A, B)                                   # the maps do not overlap
    cc, best, tab = vt.density_rotate(A, vt.VolMap(G(gA), rng_map(3, gA.n)), 30, 0, (3.5, 3.5, 3.5), -1.0)
    assert cc == -1.0 and best is None and tab.size == 0   # an empty loop leaves the running best alone
    # tsteps = 0: the grid search is the arg max of the plain pose list, first maximum wins
    pos, radii, mass = make_structure(400, 9.0, seed=91)
    gT, t = make_target(oracle, pos, radii, 5.0, 1.0, rot=(24, 48, 0), noise=0.02, seed=92, mass=mass)
    target = vt.VolMap(G(gT), t)
    com = vt.measure_center(pos, mass)
    ctx = vt.FitContext(pos, radii, target, 5.0)
    rots = np.array([[0, 0, 0], [24, 48, 0], [24, 48, 0], [48, 24, 0]], np.float32)
    bcc, bidx, bpose, allcc = ctx.search_grid(com, rots, 0, 0.0, want_all=True)
    flat = ctx.eval_poses(com, vt.FitContext.make_poses(rots))
    assert np.array_equal(bits(allcc), bits(flat)) and bidx == 1 and bcc == flat[1] and np.array_equal(bpose[:3], rots[1])
    with pytest.raises(vt.VoltoolGpuError):
        ctx.search_grid(com, np.zeros((0, 3), np.float32), 1, 1.0)
    ctx.close()
