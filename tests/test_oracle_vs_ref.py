"""Pins the C restatement (oracle/voltool_oracle.c) to the reference's OWN object code
(oracle/_ref: VolumetricData.C / Voltool.C / MDFF.C compiled unchanged).  Bit-exact everywhere
except the threaded CC sums (double summation order)."""
import numpy as np
import pytest

from oracle.pyoracle import Grid

FLT_MAX = 3.4028234663852886e38


def rng_map(seed, shape):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


def grids():
    gA = Grid.make((0.0, 0.0, 0.0), (20, 18, 16), 1.0)
    gB = Grid.make((0.5, -1.25, 2.0), (17, 21, 13), (1.1, 0.9, 1.3))
    gC = Grid.make((-3.3, 1.7, 0.4), (25, 12, 19), 0.77)
    gD = Grid.make((1.0e-3, 5.5, -2.0), (9, 31, 14), (2.0, 0.5, 1.0))
    return [gA, gB, gC, gD]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_grid(g1, g2):
    t1, t2 = g1.astuple(), g2.astuple()
    return t1[4] == t2[4] and all(np.array_equal(np.array(a), np.array(b), equal_nan=True) for a, b in zip(t1[:4], t2[:4]))


def test_voxel_coord_bit_exact(oracle, ref):
    for g in grids():
        v = ref.vol(g, np.zeros(g.n, np.float32))
        for i in list(range(0, g.n, 37)) + [g.n - 1]:
            assert bits(np.array(oracle.voxel_coord(g, i))).tolist() == bits(np.array(v.voxel_coord(i))).tolist()
            assert bits(np.array(oracle.voxel_coord_int(g, i))).tolist() == bits(np.array(v.voxel_coord_int(i))).tolist()


def test_lookups_bit_exact(oracle, ref):
    r = np.random.default_rng(7)
    for k, g in enumerate(grids()):
        d = rng_map(100 + k, g.n)
        v = ref.vol(g, d)
        lo = np.array(g.origin) - 2.0
        hi = np.array(g.origin) + np.array([g.xaxis[0], g.yaxis[1], g.zaxis[2]]) + 2.0
        pts = (lo + (hi - lo) * r.random((400, 3))).astype(np.float32)
        # exact grid points too (exercise the trunc / index-0 quirks)
        gp = np.array([oracle.voxel_coord(g, i) for i in range(0, g.n, max(1, g.n // 100))], np.float32)
        pts = np.concatenate([pts, gp, gp[:1] * 0 + np.array(g.origin, np.float32)])
        for mode in range(4):
            a, b = oracle.lookup(g, d, mode, pts), v.lookup(mode, pts)
            assert np.array_equal(bits(a), bits(b)), f"mode {mode}"


def test_intersection_union_geometry(oracle, ref):
    gs = grids()
    for A in gs:
        for B in gs:
            assert oracle.intersection(A, B).astuple() == ref.ref_intersection(A, B).astuple()
            assert oracle.union(A, B).astuple() == ref.ref_union(A, B).astuple()


@pytest.mark.parametrize("thr", [-FLT_MAX, 0.3])
def test_cc_matches_reference(oracle, ref, thr):
    gs = grids()
    ref.set_numprocs(1)
    for i, A in enumerate(gs):
        for j, B in enumerate(gs):
            a, b = rng_map(10 + i, A.n), rng_map(20 + j, B.n)
            inter = oracle.intersection(A, B)
            if inter.n == 0:
                continue
            c_ref = ref.ref_cc(A, a, B, b, thr)
            c_orc = oracle.cc(A, a, B, b, thr)
            assert (np.isnan(c_ref) and np.isnan(c_orc)) or c_ref == c_orc  # sequential: bit-exact
    ref.set_numprocs(0)


def test_cc_threaded_close(oracle, ref):
    A = Grid.make((0, 0, 0), (48, 40, 44), 1.0)
    a = rng_map(1, A.n)
    b = (0.7 * a + 0.3 * rng_map(2, A.n)).astype(np.float32)
    ref.set_numprocs(4)
    c4 = ref.ref_cc(A, a, A, b, -FLT_MAX)
    ref.set_numprocs(0)
    c1 = oracle.cc(A, a, A, b, -FLT_MAX)
    c8 = oracle.cc(A, a, A, b, -FLT_MAX, nthreads=4)
    assert abs(c4 - c1) < 1e-12 and abs(c8 - c1) < 1e-12
    assert abs(c1 - 0.7 / np.sqrt(0.58)) < 5e-3  # closed form for B = 0.7A + 0.3U


@pytest.mark.parametrize("op", [0, 1, 2, 3])
@pytest.mark.parametrize("interp", [True, False])
@pytest.mark.parametrize("use_union", [True, False])
def test_binary_ops_bit_exact(oracle, ref, op, interp, use_union):
    gs = grids()
    for i, A in enumerate(gs[:3]):
        for j, B in enumerate(gs[:3]):
            a, b = rng_map(30 + i, A.n) - 0.25, rng_map(40 + j, B.n) - 0.25
            g1, o1 = oracle.binary(op, A, a, B, b, interp, use_union)
            g2, o2 = ref.ref_binary(op, A, a, B, b, interp, use_union)
            assert g1.astuple() == g2.astuple()
            assert np.array_equal(bits(o1), bits(o2))


def test_unary_sequence_bit_exact(oracle, ref):
    g = Grid.make((1.0, 2.0, 3.0), (22, 17, 19), 1.25)
    d0 = (rng_map(5, g.n) * 4 - 1).astype(np.float32)
    seq = [("clamp", (-0.5, 2.5)), ("scale_by", (-1.5,)), ("scalar_add", (0.75,)), ("rescale_range", (0.0, 1.0)),
           ("sigma_scale", ()), ("mdff_potential", (0.2,)), ("scale_by", (3.0,)), ("binmask", (1.0,))]
    v = ref.vol(g, d0)
    d = d0.copy()
    c = oracle.new_cache(g, d)
    assert c.astuple() == v.cache.astuple()
    for name, params in seq:
        v.op(name, *params)
        d = oracle.unary(name, g, d, c, *params)
        assert np.array_equal(bits(d), bits(v.data)), name
        assert c.astuple() == v.cache.astuple(), name
    # stats after the sequence (exercises the cache-validity branches)
    assert oracle.sigma(g, d, c) == v.sigma()
    assert oracle.mean(g, d, c) == v.mean()
    assert oracle.datarange(g, d, c) == v.datarange()
    assert c.astuple() == v.cache.astuple()


def test_stats_and_com_bit_exact(oracle, ref):
    for k, g in enumerate(grids()):
        d = rng_map(60 + k, g.n) * 3 - 1
        v = ref.vol(g, d)
        c = oracle.new_cache(g, d)
        assert oracle.mean(g, d, c) == v.mean()
        assert oracle.sigma(g, d, c) == v.sigma()
        assert oracle.datarange(g, d, c) == v.datarange()
        assert np.array_equal(bits(oracle.vol_com(g, d)), bits(v.com()))
        b1, m1 = oracle.histogram(g, d, c, 10)
        b2, m2 = v.histogram(10)
        assert np.array_equal(b1, b2) and np.array_equal(bits(m1), bits(m2))


@pytest.mark.parametrize("pads", [(2, 3, 0, 1, 4, 0), (-2, -3, -1, 0, 0, -4), (3, -2, -1, 5, 0, 0), (-30, 0, 0, 0, 0, 0)])
def test_pad_bit_exact(oracle, ref, pads):
    g = Grid.make((1.0, -2.0, 0.5), (14, 11, 12), (1.0, 1.5, 0.8))
    d = rng_map(3, g.n)
    v = ref.vol(g, d).op("pad", *pads)
    g2, o = oracle.pad(g, d, pads)
    assert same_grid(g2, v.grid)
    assert np.array_equal(bits(o), bits(v.data))


def test_crop_bit_exact(oracle, ref):
    g = Grid.make((1.0, -2.0, 0.5), (14, 11, 12), (1.0, 1.5, 0.8))
    d = rng_map(4, g.n)
    for box in [(3.2, 0.1, 1.9, 9.7, 8.4, 7.7), (-3.0, -5.0, -1.0, 20.0, 20.0, 12.0), (2.0, 1.0, 1.3, 30.0, 5.0, 4.1)]:
        v = ref.vol(g, d).op("crop", *box)
        g2, o = oracle.crop(g, d, box)
        assert same_grid(g2, v.grid)
        assert np.array_equal(bits(o), bits(v.data))


@pytest.mark.parametrize("size", [(16, 12, 10), (15, 13, 11), (3, 4, 5), (7, 3, 9)])
def test_down_super_sample_bit_exact(oracle, ref, size):
    g = Grid.make((0.5, 1.0, -1.0), size, 1.2)
    d = rng_map(8, g.n) * 2 - 0.5
    v = ref.vol(g, d).op("supersample")
    g2, o = oracle.supersample(g, d)
    assert same_grid(g2, v.grid)
    assert np.array_equal(bits(o), bits(v.data))
    v = ref.vol(g, d).op("downsample")
    g3, o3 = oracle.downsample(g, d)
    assert same_grid(g3, v.grid)
    assert np.array_equal(bits(o3), bits(v.data))


def test_blur_route(oracle, ref):
    """the reference's gaussian_blur() calls the (missing) GaussianBlur class; through the glue it
    must land in the documented restatement unchanged, and invalidate the caches"""
    g = Grid.make((0, 0, 0), (20, 17, 15), 1.0)
    d = rng_map(9, g.n)
    v = ref.vol(g, d).op("blur", 1.5)
    assert np.array_equal(bits(oracle.blur(g, d, 1.5)), bits(v.data))
    assert v.cache.astuple()[:3] == (0, 0, 0)


def test_vol_move_moveto_bit_exact(oracle, ref):
    r = np.random.default_rng(12)
    for g in grids():
        d = rng_map(1, g.n)
        com, pos = r.normal(size=3).astype(np.float32) * 10, r.normal(size=3).astype(np.float32) * 10
        assert oracle.vol_moveto(g, com, pos).astuple() == ref.vol(g, d).moveto(com, pos).grid.astuple()
        a = 0.7
        mat = np.array([np.cos(a), np.sin(a), 0, 0, -np.sin(a), np.cos(a), 0, 0, 0, 0, 1, 0, 3.25, -1.5, 0.125, 1], np.float32)
        assert oracle.vol_move(g, mat).astuple() == ref.vol(g, d).move(mat).grid.astuple()


def rot_mat(axis, angle_rad):
    """Matrix4::rotate_axis for a unit axis (column-major), as the oracle restates it"""
    c, s = np.float32(np.cos(np.float64(np.float32(angle_rad)))), np.float32(np.sin(np.float64(np.float32(angle_rad))))
    m = np.eye(4, dtype=np.float32).reshape(-1)
    if axis == 0:
        m[5], m[6], m[9], m[10] = c, s, -s, c
    elif axis == 1:
        m[0], m[2], m[8], m[10] = c, -s, s, c
    else:
        m[0], m[1], m[4], m[5] = c, s, -s, c
    return m


def test_density_rotate_matches_reference_object_code(oracle, ref):
    """density_rotate (TclVoltool.C:202-279) composed from the reference's OWN vol_moveto / vol_move / vol_com /
    cc_threaded object code, candidate by candidate, against the oracle's restatement of the loop"""
    gS = Grid.make((2.0, -1.0, 0.5), (13, 12, 11), 1.5)
    gT = Grid.make((0.0, -3.0, -2.0), (24, 22, 20), 1.0)
    s, t = rng_map(5, gS.n), rng_map(6, gT.n)
    com = oracle.vol_com(gS, s)
    stride, max_rot = 25, 3
    cc_o, best_o, tab_o = oracle.density_rotate(gT, t, gS, s, stride, max_rot, com, -1.0)
    vT = ref.vol(gT, t)
    tab_r = []
    for x in range(max_rot):
        for y in range(max_rot):
            for z in range(max_rot):
                v = ref.vol(gS, s)
                v.moveto(com, (0.0, 0.0, 0.0))
                for ax, amt in enumerate((x, y, z)):
                    v.move(rot_mat(ax, np.float32(stride * amt * np.pi / 180.0)))
                v.moveto(v.com(), com)
                tab_r.append(np.float32(ref.ref_cc_vols(vT, v, 0.1)))
    tab_r = np.array(tab_r, np.float32)
    assert np.array_equal(np.isnan(tab_o), np.isnan(tab_r))
    ok = ~np.isnan(tab_r)
    assert ok.sum() >= max_rot ** 3 // 2
    assert np.max(np.abs(tab_o[ok] - tab_r[ok])) <= 1e-6      # float(cc); threaded double sums differ in order only
    best = -1.0
    bi = -1
    for i, v in enumerate(tab_r):
        if v > best:
            best, bi = v, i
    assert best_o == (bi // 9, (bi // 3) % 3, bi % 3)


def test_cc_maps_twin_is_consistent(oracle):
    """the diff / spatial-CC restatement: global value == cc_threaded's, diff == subtract on the intersection grid,
    a tile's value == the CC of that tile cropped out"""
    gs = grids()
    for k, (gA, gB) in enumerate([(gs[0], gs[0]), (gs[1], gs[0]), (gs[2], gs[3])]):
        a, b = rng_map(40 + k, gA.n), rng_map(50 + k, gB.n)
        for thr in (-FLT_MAX, 0.3):
            cc, dg, d, sg, sp = oracle.cc_maps(gA, a, gB, b, thr)
            ref_cc = oracle.cc(gA, a, gB, b, thr)
            assert (np.isnan(cc) and np.isnan(ref_cc)) or abs(cc - ref_cc) < 1e-10
            assert same_grid(dg, oracle.intersection(gA, gB))
            assert (sg.xsize, sg.ysize, sg.zsize) == ((dg.xsize + 7) // 8, (dg.ysize + 7) // 8, (dg.zsize + 7) // 8)
            assert len(sp) == sg.n and np.all(np.abs(sp[~np.isnan(sp)]) <= 1.0 + 1e-6)
        # interpolating subtract on the intersection grid is the same a - b where both samples exist
        # (subtract uses the SAFE lookups: 0 outside instead of NaN)
        og, sub = oracle.binary(1, gA, a, gB, b, interp=True, use_union=False)
        inside = ~np.isnan(d)
        assert np.array_equal(bits(sub[inside]), bits(d[inside]))
