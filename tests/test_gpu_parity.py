"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Bit-exact for elementwise ops, indexing, resampling; 1e-5 relative for sim voxels and CC
(north_star tolerance); CC of map pairs to 1e-9 (only the double summation order differs)."""
import numpy as np
import pytest

import voltool_b200 as vt
from oracle.pyoracle import Grid as OGrid
from tests.synth import make_structure, make_target

pytestmark = pytest.mark.gpu
FLT_MAX = 3.4028234663852886e38


def G(g):
    return vt.Grid.from_fields(g)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_grid(g1, g2):
    t1, t2 = g1.astuple(), g2.astuple()
    return t1[4] == t2[4] and all(np.array_equal(np.array(a), np.array(b), equal_nan=True) for a, b in zip(t1[:4], t2[:4]))


def same_grid_tuple(g1, g2, tol=1e-5):
    """geometry after a trip through MRC's float header words"""
    t1, t2 = g1.astuple(), g2.astuple()
    return t1[4] == t2[4] and all(np.allclose(np.array(a), np.array(b), rtol=tol, atol=tol) for a, b in zip(t1[:4], t2[:4]))


def rng_map(seed, n):
    return np.random.default_rng(seed).random(n, dtype=np.float32)


def grids():
    return [OGrid.make((0.0, 0.0, 0.0), (20, 18, 16), 1.0),
            OGrid.make((0.5, -1.25, 2.0), (17, 21, 13), (1.1, 0.9, 1.3)),
            OGrid.make((-3.3, 1.7, 0.4), (25, 12, 19), 0.77),
            OGrid.make((1.0e-3, 5.5, -2.0), (9, 31, 14), (2.0, 0.5, 1.0))]


def skewed_grid():
    # non-orthogonal cell: forces the generic (per-voxel matrix) path
    return OGrid.make((0.3, 0.1, -0.2), (15, 14, 13), axes=((14.0, 1.0, 0.0), (0.5, 13.0, 0.7), (0.0, 0.3, 12.0)))


@pytest.fixture(scope="module", autouse=True)
def _init():
    vt.init(0)
    yield


# ------------------------------------------------------------------------------------------ cc
@pytest.mark.parametrize("thr", [-FLT_MAX, 0.3])
def test_cc_pairs(oracle, thr):
    gs = grids()
    for i, A in enumerate(gs):
        for j, B in enumerate(gs):
            a, b = rng_map(10 + i, A.n), rng_map(20 + j, B.n)
            if oracle.intersection(A, B).n <= 0:
                continue
            want, nw, sw = oracle.cc(A, a, B, b, thr, full=True)
            got, ng, sg = vt.cc_threaded(vt.VolMap(G(A), a), vt.VolMap(G(B), b), thr, full=True)
            assert ng == nw
            if np.isnan(want):
                assert np.isnan(got)
            else:
                assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (i, j, got, want)
                assert np.allclose(sg, sw, rtol=1e-6, atol=1e-6)


def test_cc_generic_path_matches(oracle, monkeypatch):
    A, B = grids()[0], grids()[1]
    a, b = rng_map(1, A.n), rng_map(2, B.n)
    want = oracle.cc(A, a, B, b, 0.2)
    monkeypatch.setenv("VOLTOOL_FORCE_GENERIC", "1")
    got = vt.cc_threaded(vt.VolMap(G(A), a), vt.VolMap(G(B), b), 0.2)
    assert abs(got - want) <= 1e-12
    # skewed cells take the generic path on their own
    monkeypatch.delenv("VOLTOOL_FORCE_GENERIC")
    S = skewed_grid()
    s = rng_map(3, S.n)
    want = oracle.cc(S, s, A, a, -FLT_MAX)
    got = vt.cc_threaded(vt.VolMap(G(S), s), vt.VolMap(G(A), a), -FLT_MAX)
    assert abs(got - want) <= 1e-9


def test_cc_staged_ring_path(oracle, monkeypatch):
    """the cp.async.bulk / mbarrier ring variant (forced here on small maps): same samples, sums formed
    in float per 8 voxels -> 1e-7 on CC; falls back to the direct kernel when its preconditions fail"""
    r = np.random.default_rng(11)
    cases = [(OGrid.make((0, 0, 0), (96, 70, 150), 1.0), OGrid.make((0, 0, 0), (96, 70, 150), 1.0)),      # identical
             (OGrid.make((0, 0, 0), (128, 64, 90), 1.0), OGrid.make((0.5, -3.0, 2.25), (132, 60, 96), 1.0)),  # offset
             (OGrid.make((1.0, 2.0, 3.0), (80, 81, 79), 1.5), OGrid.make((0, 0, 0), (128, 128, 128), 1.0)),  # fit-like
             (OGrid.make((0, 0, 0), (100, 40, 40), 1.0), OGrid.make((0.3, 0.2, 0.1), (60, 44, 48), 0.9))]   # nx % 4 != 0
    for A, B in cases:
        a, b = r.random(A.n, dtype=np.float32), r.random(B.n, dtype=np.float32)
        b[::1013] = np.nan
        for thr in (-FLT_MAX, 0.4):
            want, nw, _ = oracle.cc(A, a, B, b, thr, full=True)
            monkeypatch.setenv("VOLTOOL_CC_STAGED", "1")
            got, ng, _ = vt.cc_threaded(vt.VolMap(G(A), a), vt.VolMap(G(B), b), thr, full=True)
            monkeypatch.setenv("VOLTOOL_CC_STAGED", "0")
            ref, nr, _ = vt.cc_threaded(vt.VolMap(G(A), a), vt.VolMap(G(B), b), thr, full=True)
            monkeypatch.delenv("VOLTOOL_CC_STAGED")
            assert ng == nw == nr
            assert abs(ref - want) <= 1e-9 and abs(got - want) <= 1e-7, (got, ref, want)


def test_cc_identity_stream_path(oracle, monkeypatch):
    """maps on one exactly representable lattice (whole-voxel shifts allowed) are streamed once without
    interpolation; a NaN/inf/huge voxel anywhere sends the call back to the interpolating path"""
    r = np.random.default_rng(12)
    cases = [(OGrid.make((0, 0, 0), (64, 60, 56), 1.0), OGrid.make((0, 0, 0), (64, 60, 56), 1.0)),        # identical
             (OGrid.make((0, 0, 0), (40, 36, 32), 1.0), OGrid.make((-4, -8, -2), (64, 60, 56), 1.0)),     # aligned sub-box
             (OGrid.make((0, 0, 0), (40, 36, 32), 0.5), OGrid.make((-1.5, -4, -1), (64, 60, 56), 0.5)),   # x shift 3: scalar loads
             (OGrid.make((2, 2, 2), (37, 21, 19), 2.0), OGrid.make((2, 2, 2), (37, 21, 19), 2.0))]        # odd sizes
    for A, B in cases:
        a, b = r.random(A.n, dtype=np.float32), r.random(B.n, dtype=np.float32)
        for thr in (-FLT_MAX, 0.4):
            want, nw, sw = oracle.cc(A, a, B, b, thr, full=True)
            monkeypatch.setenv("VOLTOOL_CC_IDENTITY", "1")
            got, ng, sg = vt.cc_threaded(vt.VolMap(G(A), a), vt.VolMap(G(B), b), thr, full=True)
            monkeypatch.setenv("VOLTOOL_CC_IDENTITY", "0")
            ref, nr, _ = vt.cc_threaded(vt.VolMap(G(A), a), vt.VolMap(G(B), b), thr, full=True)
            monkeypatch.delenv("VOLTOOL_CC_IDENTITY")
            assert ng == nw == nr
            assert abs(ref - want) <= 1e-9 and abs(got - want) <= 1e-7, (got, ref, want)
            assert np.allclose(sg, sw, rtol=1e-6)
        # values that break "sample == voxel": the answer must still be the oracle's
        for poison in (np.nan, np.inf, -3.0e38):
            for which in (0, 1):
                a2, b2 = a.copy(), b.copy()
                (a2 if which == 0 else b2)[::997] = poison
                want, nw, _ = oracle.cc(A, a2, B, b2, 0.2, full=True)
                monkeypatch.setenv("VOLTOOL_CC_IDENTITY", "1")
                got, ng, _ = vt.cc_threaded(vt.VolMap(G(A), a2), vt.VolMap(G(B), b2), 0.2, full=True)
                monkeypatch.delenv("VOLTOOL_CC_IDENTITY")
                assert ng == nw
                if np.isnan(want):
                    assert np.isnan(got)
                else:
                    assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (poison, which, got, want)


def test_cc_identity_poison_bordering_the_overlap(oracle, monkeypatch):
    """the reference's sample of a box voxel reads its +1 neighbours: a NaN / inf / huge voxel of the LARGER map just
    beyond the common box's upper faces must reach the result although the streamed box never contains it"""
    A = OGrid.make((0, 0, 0), (40, 36, 32), 1.0)
    B = OGrid.make((-4, -8, -2), (64, 60, 56), 1.0)         # A is an aligned sub-box of B, ending inside it
    r = np.random.default_rng(14)
    a, b = r.random(A.n, dtype=np.float32), r.random(B.n, dtype=np.float32)
    og = oracle.intersection(A, B)
    ox, oy, oz = og.xsize, og.ysize, og.zsize
    bz, by, bx = B.shape
    b3 = b.reshape(bz, by, bx)
    monkeypatch.setenv("VOLTOOL_CC_IDENTITY", "1")
    for poison in (np.nan, np.inf, 3.0e38):
        for where in ((2 + oz, 8 + 5, 4 + 7), (2 + 3, 8 + oy, 4 + 9), (2 + 6, 8 + 11, 4 + ox), (2 + oz, 8 + oy, 4 + ox)):
            b2 = b3.copy()
            b2[where] = poison                                # one voxel past the box on the +z / +y / +x face / the corner
            want, nw, _ = oracle.cc(A, a, B, b2.reshape(-1), -FLT_MAX, full=True)
            got, ng, _ = vt.cc_threaded(vt.VolMap(G(A), a), vt.VolMap(G(B), b2.reshape(-1)), -FLT_MAX, full=True)
            assert ng == nw, (poison, where)
            assert (np.isnan(want) and np.isnan(got)) or abs(got - want) <= 1e-9 * max(1.0, abs(want)), (poison, where, got, want)
            for op in (0, 2):                                 # and the two-map ops (bit-exact)
                gw, want_o = oracle.binary(op, A, a, B, b2.reshape(-1), True, False)
                got_o = vt.binary(op, vt.VolMap(G(A), a), vt.VolMap(G(B), b2.reshape(-1)), True, False)
                assert np.array_equal(bits(got_o.data), bits(want_o)), (op, poison, where)


def test_cc_identity_low_variance_tolerance(oracle, monkeypatch):
    """the streaming kernel adds each group of four voxel terms in float before promoting to double (MDFF.C:79-83
    promotes every term); on maps whose variance is small against their mean the cancellation in
    sumAA/n - mean^2 amplifies that.  Stated tolerance: 1e-5 on CC (north star) down to sd/mean = 1e-2."""
    g = OGrid.make((0, 0, 0), (64, 60, 56), 1.0)
    r = np.random.default_rng(15)
    for rel in (1e-1, 1e-2):
        a = (10.0 * (1.0 + rel * r.standard_normal(g.n))).astype(np.float32)
        b = (0.6 * a + 4.0 * (1.0 + rel * r.standard_normal(g.n))).astype(np.float32)
        want = oracle.cc(g, a, g, b)
        monkeypatch.setenv("VOLTOOL_CC_IDENTITY", "1")
        got = vt.cc_threaded(vt.VolMap(G(g), a), vt.VolMap(G(g), b))
        monkeypatch.setenv("VOLTOOL_CC_IDENTITY", "0")
        ref = vt.cc_threaded(vt.VolMap(G(g), a), vt.VolMap(G(g), b))
        monkeypatch.delenv("VOLTOOL_CC_IDENTITY")
        assert abs(ref - want) <= 1e-9
        assert abs(got - want) <= 1e-5 * abs(want), (rel, got, want)


def test_cc_known_answers():
    g = vt.Grid.make((0, 0, 0), (64, 60, 56), 1.0)
    a = rng_map(5, g.n)
    u = rng_map(6, g.n)
    A = vt.VolMap(g, a)
    assert abs(vt.cc_threaded(A, A) - 1.0) < 1e-9
    assert abs(vt.cc_threaded(A, vt.VolMap(g, -a)) + 1.0) < 1e-9
    b = (0.7 * a + 0.3 * u).astype(np.float32)
    assert abs(vt.cc_threaded(A, vt.VolMap(g, b)) - 0.7 / np.sqrt(0.58)) < 5e-3
    # invariance to a positive affine rescale of either map
    c1 = vt.cc_threaded(A, vt.VolMap(g, b))
    c2 = vt.cc_threaded(A, vt.VolMap(g, (3.0 * b + 1.0).astype(np.float32)))
    assert abs(c1 - c2) < 1e-5


def test_cc_nan_and_threshold(oracle):
    A, B = grids()[0], grids()[2]
    a, b = rng_map(7, A.n), rng_map(8, B.n)
    a[::17] = np.nan
    b[::29] = np.nan
    for thr in (-FLT_MAX, 0.5):
        want, nw, _ = oracle.cc(A, a, B, b, thr, full=True)
        got, ng, _ = vt.cc_threaded(vt.VolMap(G(A), a), vt.VolMap(G(B), b), thr, full=True)
        assert ng == nw and abs(got - want) <= 1e-9


# ------------------------------------------------------------------------------------------ two-map ops
@pytest.mark.parametrize("op", [0, 1, 2, 3])
@pytest.mark.parametrize("interp", [True, False])
@pytest.mark.parametrize("use_union", [True, False])
def test_binary_ops_bit_exact(oracle, op, interp, use_union):
    gs = grids()[:3] + [skewed_grid()]
    for i, A in enumerate(gs):
        for j, B in enumerate(gs):
            if use_union and (i == 3 or j == 3):
                continue  # init_from_union is orthogonal-only (Voltool.C:118-133)
            a, b = rng_map(30 + i, A.n) - 0.25, rng_map(40 + j, B.n) - 0.25
            og, want = oracle.binary(op, A, a, B, b, interp, use_union)
            if og.n <= 0:
                continue
            got = vt.binary(op, vt.VolMap(G(A), a), vt.VolMap(G(B), b), interp, use_union)
            assert got.grid.astuple() == og.astuple()
            assert np.array_equal(bits(got.data), bits(want)), (op, interp, use_union, i, j)


def test_binary_ops_staged_ring_path_bit_exact(oracle, monkeypatch):
    """interpolated two-map ops through the cp.async.bulk ring (forced on small maps): bit-exact,
    including voxels that involve NaNs (redone with the literal arithmetic) and planes outside a map"""
    r = np.random.default_rng(21)
    cases = [(OGrid.make((0, 0, 0), (96, 70, 50), 1.0), OGrid.make((0, 0, 0), (96, 70, 50), 1.0)),
             (OGrid.make((0, 0, 0), (128, 64, 60), 1.0), OGrid.make((0.5, -3.0, 2.25), (132, 60, 66), 1.0)),
             (OGrid.make((1.0, 2.0, 3.0), (40, 41, 39), 1.5), OGrid.make((0, 0, 0), (64, 64, 64), 1.0))]
    monkeypatch.setenv("VOLTOOL_CC_STAGED", "1")
    for A, B in cases:
        a, b = r.random(A.n, dtype=np.float32) - 0.3, r.random(B.n, dtype=np.float32) - 0.3
        # NaNs in the first half of A, infs (-> generated NaNs) in the second half of B: a voxel where BOTH
        # operands are NaNs of different payload is left out (there x86 returns whichever operand the
        # compiler put first)
        a[:A.n // 3:997] = np.nan
        b[2 * B.n // 3 + 5::1499] = np.inf
        for op in range(4):
            for use_union in (False, True):
                og, want = oracle.binary(op, A, a, B, b, True, use_union)
                got = vt.binary(op, vt.VolMap(G(A), a), vt.VolMap(G(B), b), True, use_union)
                assert got.grid.astuple() == og.astuple()
                assert np.array_equal(bits(got.data), bits(want)), (op, use_union)


def test_binary_ops_identity_stream_path_bit_exact(oracle, monkeypatch):
    """two-map ops on one exact lattice stream without interpolation; -0 voxels (whose sampled sign
    depends on their neighbours) are redone literally, NaN/inf/huge values send the op back to the
    interpolating path: every result bit-exact"""
    r = np.random.default_rng(22)
    cases = [(OGrid.make((0, 0, 0), (96, 70, 50), 1.0), OGrid.make((0, 0, 0), (96, 70, 50), 1.0)),
             (OGrid.make((0, 0, 0), (40, 36, 32), 1.0), OGrid.make((-4, -8, -2), (64, 60, 56), 1.0)),
             (OGrid.make((2, 2, 2), (37, 21, 19), 2.0), OGrid.make((2, 2, 2), (37, 21, 19), 2.0))]
    monkeypatch.setenv("VOLTOOL_CC_STAGED", "1")
    for A, B in cases:
        a, b = r.random(A.n, dtype=np.float32) - 0.3, r.random(B.n, dtype=np.float32) - 0.3
        a[::7] = 0.0
        a[::11] = -0.0
        b[::13] = -0.0
        b[5::17] = 0.0
        for poison in (None, np.nan, np.inf, 3.0e38):
            a2 = a.copy()
            if poison is not None:
                a2[3::1009] = poison
            for op in range(4):
                for use_union in (False, True):
                    og, want = oracle.binary(op, A, a2, B, b, True, use_union)
                    got = vt.binary(op, vt.VolMap(G(A), a2), vt.VolMap(G(B), b), True, use_union)
                    assert got.grid.astuple() == og.astuple()
                    assert np.array_equal(bits(got.data), bits(want)), (op, use_union, poison)


# ------------------------------------------------------------------------------------------ unary + stats
def test_unary_sequence_bit_exact(oracle):
    g = OGrid.make((1.0, 2.0, 3.0), (22, 17, 19), 1.25)
    d0 = (rng_map(5, g.n) * 4 - 1).astype(np.float32)
    seq = [("clamp", (-0.5, 2.5)), ("scale_by", (-1.5,)), ("scalar_add", (0.75,)), ("rescale_range", (0.0, 1.0)),
           ("sigma_scale", ()), ("mdff_potential", (0.2,)), ("scale_by", (3.0,)), ("binmask", (1.0,))]
    names = {"rescale_range": "rescale_voxel_value_range"}
    m = vt.VolMap(G(g), d0)
    d = d0.copy()
    c = oracle.new_cache(g, d)
    assert c.astuple() == m.cache.astuple()
    for name, params in seq:
        getattr(m, names.get(name, name))(*params)
        d = oracle.unary(name, g, d, c, *params)
        assert np.array_equal(bits(m.data), bits(d)), name
        assert c.astuple() == m.cache.astuple(), name
    assert oracle.sigma(g, d, c) == m.sigma()
    assert oracle.mean(g, d, c) == m.mean()
    assert oracle.datarange(g, d, c) == m.datarange()
    assert c.astuple() == m.cache.astuple()


def test_stats(oracle):
    for k, g in enumerate(grids()):
        d = rng_map(60 + k, g.n) * 3 - 1
        m = vt.VolMap(G(g), d)
        c = oracle.new_cache(g, d)
        # double sums in a different order: identical after the float cast except on a rounding tie
        assert abs(m.mean() - oracle.mean(g, d, c)) <= 1.2e-7 * abs(oracle.mean(g, d, c))
        assert abs(m.sigma() - oracle.sigma(g, d, c)) <= 1.2e-7 * abs(oracle.sigma(g, d, c))
        assert m.datarange() == oracle.datarange(g, d, c)
        b1, m1 = oracle.histogram(g, d, c, 10)
        b2, m2 = m.histogram(10)
        assert np.array_equal(b1, b2) and np.array_equal(bits(m1), bits(m2))


def test_vol_com_bit_exact(oracle):
    for k, g in enumerate(grids()):
        d = rng_map(70 + k, g.n) * 3 - 0.5
        assert np.array_equal(bits(vt.VolMap(G(g), d).vol_com()), bits(oracle.vol_com(g, d)))


def test_unaligned_sizes_and_tiny_maps(oracle):
    for size in [(1, 1, 1), (3, 1, 2), (5, 7, 3), (33, 2, 1)]:
        g = OGrid.make((0, 0, 0), size, 1.0)
        d = rng_map(9, g.n) - 0.5
        m = vt.VolMap(G(g), d)
        c = oracle.new_cache(g, d)
        assert m.datarange() == oracle.datarange(g, d, c)
        m.scale_by(2.5).scalar_add(-1.0).binmask(0.1)
        e = oracle.unary("scale_by", g, d, c, 2.5)
        e = oracle.unary("scalar_add", g, e, c, -1.0)
        e = oracle.unary("binmask", g, e, c, 0.1)
        assert np.array_equal(bits(m.data), bits(e))


# ------------------------------------------------------------------------------------------ resizing
@pytest.mark.parametrize("pads", [(2, 3, 0, 1, 4, 0), (-2, -3, -1, 0, 0, -4), (3, -2, -1, 5, 0, 0), (-30, 0, 0, 0, 0, 0)])
def test_pad_bit_exact(oracle, pads):
    g = OGrid.make((1.0, -2.0, 0.5), (14, 11, 12), (1.0, 1.5, 0.8))
    d = rng_map(3, g.n)
    g2, want = oracle.pad(g, d, pads)
    m = vt.VolMap(G(g), d).pad(*pads)
    assert m.grid.astuple() == g2.astuple()
    assert np.array_equal(bits(m.data), bits(want))
    assert m.cache.astuple()[:3] == (0, 0, 0)


def test_crop_bit_exact(oracle):
    g = OGrid.make((1.0, -2.0, 0.5), (14, 11, 12), (1.0, 1.5, 0.8))
    d = rng_map(4, g.n)
    for box in [(3.2, 0.1, 1.9, 9.7, 8.4, 7.7), (-3.0, -5.0, -1.0, 20.0, 20.0, 12.0), (2.0, 1.0, 1.3, 30.0, 5.0, 4.1)]:
        g2, want = oracle.crop(g, d, box)
        m = vt.VolMap(G(g), d).crop(*box)
        assert m.grid.astuple() == g2.astuple()
        assert np.array_equal(bits(m.data), bits(want))


@pytest.mark.parametrize("size", [(16, 12, 10), (15, 13, 11), (3, 4, 5), (7, 3, 9), (40, 33, 21)])
def test_down_super_sample_bit_exact(oracle, size):
    g = OGrid.make((0.5, 1.0, -1.0), size, 1.2)
    d = rng_map(8, g.n) * 2 - 0.5
    g2, want = oracle.supersample(g, d)
    m = vt.VolMap(G(g), d).supersample()
    assert same_grid(m.grid, g2)
    assert np.array_equal(bits(m.data), bits(want))
    g3, want3 = oracle.downsample(g, d)
    m3 = vt.VolMap(G(g), d).downsample()
    assert same_grid(m3.grid, g3)
    assert np.array_equal(bits(m3.data), bits(want3))


@pytest.mark.parametrize("mode", ["1", "2"])
def test_supersample_single_kernel_bit_exact(oracle, monkeypatch, mode):
    """the one-kernel forms (2 = register z window, the default for large maps; 1 = shared-memory
    ring) against the oracle on ragged sizes"""
    monkeypatch.setenv("VOLTOOL_SUPERSAMPLE_FUSED", mode)
    for size in [(70, 9, 67), (3, 3, 3), (65, 17, 130), (129, 8, 5), (64, 8, 64), (66, 10, 4), (4, 200, 3)]:
        g = OGrid.make((0.5, 1.0, -2.0), size, (1.0, 1.5, 0.5))
        a = rng_map(61, g.n)
        og, want = oracle.supersample(g, a)
        M = vt.VolMap(G(g), a)
        M.supersample()
        assert same_grid(M.grid, og) and np.array_equal(bits(M.data), bits(want)), size


def test_supersample_packed_kernel_tiny_values(oracle, monkeypatch):
    """the packed kernel multiplies by exact powers of two inside fused multiply-adds; that only differs
    from the reference's separate roundings when a product is subnormal, which it detects (a nonzero
    |voxel| < 2^-100) and hands to the scalar kernel: subnormal-range maps stay bit-exact too"""
    monkeypatch.setenv("VOLTOOL_SUPERSAMPLE_FUSED", "2")
    g = OGrid.make((0, 0, 0), (20, 18, 9), 1.0)
    rng = np.random.default_rng(77)
    a = (rng.random(g.n, dtype=np.float32) * np.float32(3e-38)).astype(np.float32)
    a[::7] = np.float32(1.0e-42)
    a[::11] = 0.0
    og, want = oracle.supersample(g, a)
    M = vt.VolMap(G(g), a)
    M.supersample()
    assert same_grid(M.grid, og) and np.array_equal(bits(M.data), bits(want))
    b = rng.random(g.n, dtype=np.float32)
    b[5] = np.float32(1.0e-35)   # one tiny voxel in an ordinary map
    og, want = oracle.supersample(g, b)
    M = vt.VolMap(G(g), b)
    M.supersample()
    assert np.array_equal(bits(M.data), bits(want))


def test_supersample_known_answers():
    # (-y0 + 5 y1 + 5 y2 - y3)/8 interior stencil; replicated ends give 3/8 on the ramp 0,1,2
    g = vt.Grid.make((0, 0, 0), (6, 3, 3), 1.0)
    ramp = np.tile(np.arange(6, dtype=np.float32), 9)
    out = vt.VolMap(g, ramp).supersample().data.reshape(5, 5, 11)
    assert out[0, 0, 3] == 1.5 and out[0, 0, 5] == 2.5
    assert out[0, 0, 1] == 0.375
    const = vt.VolMap(g, np.full(g.n, 2.5, np.float32)).downsample().data
    assert np.all(const == 2.5)


@pytest.mark.parametrize("mode", ["", "yz", "xyz", "tma", "tma_plain"])
@pytest.mark.parametrize("sigma", [0.3, 0.6, 1.0, 1.5, 2.0, 2.5])
def test_blur_bit_exact(oracle, sigma, mode, monkeypatch):
    """every form of the blur (two-kernel default, x | y+z with the z pass as a register scatter, the cp.async
    one-kernel form, the tensor-map TMA one-pass form and its plain-load twin), incl. maps thinner than the radius,
    several tiles per axis and ragged tile edges"""
    if mode:
        monkeypatch.setenv("VOLTOOL_BLUR_MODE", mode[:3])
        if mode == "tma_plain":      # same kernel, the planes staged by plain clamped loads instead of TMA
            monkeypatch.setenv("VOLTOOL_BLUR_NOTMA", "1")
    for size in [(20, 17, 15), (64, 9, 33), (5, 4, 3), (36, 35, 70), (8, 3, 2), (132, 40, 5), (48, 130, 40), (100, 70, 90)]:
        g = OGrid.make((0, 0, 0), size, 1.0)
        d = rng_map(9, g.n)
        m = vt.VolMap(G(g), d).gaussian_blur(sigma)
        assert np.array_equal(bits(m.data), bits(oracle.blur(g, d, sigma)))
        assert m.cache.astuple()[:3] == (0, 0, 0)


# ------------------------------------------------------------------------------------------ sim
def rel_err(got, want):
    return np.abs(got - want) / np.maximum(np.abs(want), 1e-30)


@pytest.mark.parametrize("res,spacing,natoms", [(5.0, 1.0, 1500), (5.0, 0.0, 2000), (3.0, 0.0, 800), (10.0, 0.0, 600)])
def test_sim_matches_oracle(oracle, res, spacing, natoms):
    pos, radii, _ = make_structure(natoms, 14.0, seed=int(res * 10) + natoms)
    radscale, gsp, quality = oracle.sim_policy(res, spacing)
    g, want = oracle.sim(pos, radii, quality, radscale, gsp, nthreads=4)
    m = vt.sim(pos, radii, quality, radscale, gsp)
    assert m.grid.astuple() == g.astuple()
    got = m.data
    # identical support (the cutoff decisions are reproduced exactly) ...
    assert np.array_equal(got == 0, want == 0)
    # ... and 1e-5 relative on every voxel
    assert rel_err(got, want).max() < 1e-5
    mn, mx = m.datarange()
    assert mn == got.min() and mx == got.max()


def test_sim_single_atom_is_sampled_gaussian():
    pos = np.array([[10.0, 11.0, 12.0]], np.float32)
    radii = np.array([1.5], np.float32)
    m = vt.sim(pos, radii, 3, 1.0, 0.5)
    g = m.grid
    d = m.data.reshape(g.zsize, g.ysize, g.xsize)
    z, y, x = np.unravel_index(np.argmax(d), d.shape)
    sig = 1.5
    for dx in (0, 1, 3):
        r2 = sum((g.origin[k] + idx * 0.5 - pos[0, k]) ** 2 for k, idx in enumerate((x + dx, y, z)))
        assert abs(d[z, y, x + dx] - np.exp(-r2 / (2 * sig * sig))) < 1e-5


# ------------------------------------------------------------------------------------------ calc_cc / fit
def test_calc_cc_and_pose_batch(oracle):
    pos, radii, mass = make_structure(1200, 13.0, seed=21)
    g, t = make_target(oracle, pos, radii, 5.0, 1.0, rot=(24, 48, 0), noise=0.02, seed=22, mass=mass)
    target = vt.VolMap(G(g), t)
    want = np.float32(oracle.calc_cc(pos, radii, g, t, 5.0))
    got = vt.calc_cc(pos, radii, target, 5.0)
    assert abs(got - want) <= 1e-5 * abs(want)
    com = oracle.measure_center(pos, mass)
    assert np.array_equal(bits(vt.measure_center(pos, mass)), bits(com))
    ctx = vt.FitContext(pos, radii, target, 5.0)
    rots = np.array([[0, 0, 0], [24, 48, 0], [48, 24, 312], [336, 0, 24], [120, 240, 96]], np.float32)
    cc = ctx.eval_poses(com, vt.FitContext.make_poses(rots))
    for r, c in zip(rots, cc):
        p = oracle.pose_transform(pos, com, 24, int(r[0]) // 24, int(r[1]) // 24, int(r[2]) // 24)
        assert np.array_equal(bits(vt.pose_apply(pos, com, r)), bits(p))
        w = np.float32(oracle.calc_cc(p, radii, g, t, 5.0))
        assert abs(float(c) - float(w)) <= 1e-5 * abs(float(w)), (r, c, w)
    # the planted pose scores best
    assert np.argmax(cc) == 1
    # batched == one-at-a-time (the batch path and the single path share no host code)
    p = vt.pose_apply(pos, com, rots[2])
    assert abs(vt.calc_cc(p, radii, target, 5.0) - float(cc[2])) <= 2e-6


def test_rotate_table_matches_oracle_stage(oracle):
    pos, radii, mass = make_structure(500, 10.0, seed=31)
    g, t = make_target(oracle, pos, radii, 5.0, 1.0, rot=(4, 8, 20), noise=0.02, seed=32, mass=mass)
    com = oracle.measure_center(pos, mass)
    want = oracle.rotate_table(pos, radii, com, 4, 6, g, t, 5.0, nthreads=8)
    ctx = vt.FitContext(pos, radii, vt.VolMap(G(g), t), 5.0)
    got = ctx.rotate_table(com, 4, 6)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-5
    cw, bw, vw = oracle.rotate_replay(want, 6, -1.0)
    cg, bg, vg = vt.rotate_replay(got, 6, -1.0)
    assert bw == bg and vw == vg and abs(cw - cg) < 1e-5


def test_fit_matches_oracle(oracle):
    pos, radii, mass = make_structure(400, 9.0, seed=41)
    g, t = make_target(oracle, pos, radii, 6.0, 1.2, rot=(48, 24, 312), noise=0.01, seed=42, mass=mass, pad=6)
    # displace the structure so the COM alignment has something to do
    start = pos + np.array([3.0, -2.0, 1.5], np.float32)
    # the LITERAL sequential loop of the reference (calc_cc per visited candidate, skip rule, stale
    # framepos after reset_origin); one thread so the double sums are in voxel order
    p_want, r_want = oracle.fit(start, radii, mass, g, t, 6.0, direct=True, nthreads=1)
    p_got, r_got = vt.fit(start, radii, mass, vt.VolMap(G(g), t), 6.0)
    assert list(r_got.stage_best) == list(r_want.stage_best)
    assert r_got.n_evaluated == r_want.n_evaluated
    assert abs(r_got.best_cc - r_want.best_cc) <= 1e-5 * abs(r_want.best_cc)
    assert np.array_equal(bits(np.array(r_got.dencom)), bits(np.array(r_want.dencom)))
    assert np.array_equal(bits(p_got), bits(p_want.reshape(-1, 3)))


def test_vol_move_moveto(oracle):
    r = np.random.default_rng(12)
    for g in grids():
        d = rng_map(1, g.n)
        com, pos = r.normal(size=3).astype(np.float32) * 10, r.normal(size=3).astype(np.float32) * 10
        assert vt.VolMap(G(g), d).vol_moveto(com, pos).grid.astuple() == oracle.vol_moveto(g, com, pos).astuple()
        a = 0.7
        mat = np.array([np.cos(a), np.sin(a), 0, 0, -np.sin(a), np.cos(a), 0, 0, 0, 0, 1, 0, 3.25, -1.5, 0.125, 1], np.float32)
        m = vt.VolMap(G(g), d).vol_move(mat)
        assert m.grid.astuple() == oracle.vol_move(g, mat).astuple()
        assert np.array_equal(bits(m.data), bits(d))


def test_map_dx_files(tmp_path):
    g = vt.Grid.make((0.5, -2.0, 3.0), (12, 9, 7), (1.0, 1.5, 0.75))
    a = rng_map(31, g.n)
    A = vt.VolMap(g, a)
    A.gaussian_blur(1.0)
    p = str(tmp_path / "blur.dx")
    A.write_dx(p, "smoothed")
    B = vt.VolMap.read_dx(p)
    assert np.array_equal(bits(A.data), bits(B.data))
    assert A.datarange() == B.datarange()          # the min/max cache is rebuilt on load
    assert abs(vt.cc_threaded(A, B) - 1.0) < 1e-9


def test_cli_commands_match_api(tmp_path, capsys):
    """the Tcl-free front end (python -m voltool_b200 ...) drives the same entry points"""
    from voltool_b200 import cli
    g = vt.Grid.make((0, 0, 0), (24, 20, 16), 1.0)
    a, b = rng_map(41, g.n), rng_map(42, g.n)
    pa, pb, po = (str(tmp_path / n) for n in ("a.dx", "b.dx", "o.dx"))
    vt.write_dx(pa, g, a)
    vt.write_dx(pb, g, b)
    assert cli.main(["smooth", "-sigma", "1.5", "-i", pa, "-o", po]) == 0
    want = vt.VolMap(g, a).gaussian_blur(1.5).data
    assert np.array_equal(bits(vt.read_dx(po)[1]), bits(want))
    assert cli.main(["trim", "-amt", "1 2 0 1 3 0", "-i", pa, "-o", po]) == 0
    T = vt.VolMap(g, a); T.pad(-1, -2, 0, -1, -3, 0)
    g2, d2 = vt.read_dx(po)
    assert (g2.xsize, g2.ysize, g2.zsize) == (T.grid.xsize, T.grid.ysize, T.grid.zsize) and np.array_equal(bits(d2), bits(T.data))
    assert cli.main(["add", "-i1", pa, "-i2", pb, "-o", po]) == 0
    assert np.array_equal(bits(vt.read_dx(po)[1]), bits(vt.add(vt.VolMap(g, a), vt.VolMap(g, b)).data))
    capsys.readouterr()
    assert cli.main(["correlate", "-i1", pa, "-i2", pb]) == 0
    assert abs(float(capsys.readouterr().out) - vt.cc_threaded(vt.VolMap(g, a), vt.VolMap(g, b))) < 1e-8
    # atoms: sim -> cc against its own map is 1
    pos, radii, mass = make_structure(300, 8.0, seed=7)
    pt, ps = str(tmp_path / "atoms.txt"), str(tmp_path / "sim.dx")
    cli.write_atoms(pt, pos, radii, mass)
    assert cli.main(["sim", "-atoms", pt, "-res", "5", "-o", ps]) == 0
    capsys.readouterr()
    assert cli.main(["cc", "-atoms", pt, "-res", "5", "-i", ps]) == 0
    assert abs(float(capsys.readouterr().out) - 1.0) < 1e-6
    assert cli.main(["smooth", "-i", pa]) == 1          # -sigma missing: usage error, nothing written
    # MRC / CCP4 maps through the same commands (chosen by extension)
    pm, pom = str(tmp_path / "a.mrc"), str(tmp_path / "o.ccp4")
    vt.VolMap(g, a).write_mrc(pm)
    M = vt.VolMap.read_mrc(pm)
    assert np.array_equal(bits(M.data), bits(a)) and same_grid_tuple(M.grid, g)
    assert cli.main(["downsample", "-i", pm, "-o", pom]) == 0
    D = vt.VolMap(g, a).downsample()
    g3, d3 = vt.read_mrc(pom)
    assert (g3.xsize, g3.ysize, g3.zsize) == (D.grid.xsize, D.grid.ysize, D.grid.zsize)
    assert np.array_equal(bits(d3), bits(D.data))
