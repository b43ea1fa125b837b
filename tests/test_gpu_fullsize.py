"""GPU tests at BASELINE-scale sizes, where the CPU oracle would take minutes: size-independent
properties of the operators (copies that must survive a resample bit for bit, idempotence, linearity of
the CC sums, CC(A, A) = 1, a checksum of a map against the checksum of its pieces), and the
equivalence of the two-stream pose pipeline with the serial one."""
import os

import numpy as np
import pytest

import voltool_b200 as vt
from voltool_b200.synth import make_structure, add_noise

pytestmark = pytest.mark.gpu
FLT_MAX = 3.4028234663852886e38


@pytest.fixture(scope="module", autouse=True)
def _init():
    vt.init(0)
    yield


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def big_map(n, seed, origin=(0, 0, 0)):
    g = vt.Grid.make(origin, (n, n, n), 1.0)
    return g, np.random.default_rng(seed).random(g.n, dtype=np.float32)


def test_supersample_large_copies_and_symmetry():
    """384^3 -> 767^3 through the packed one-kernel path: even output voxels are the source voxels bit for
    bit (VolumetricData.C:783-794); a map that is constant along an axis stays constant along it"""
    n = 384
    g, a = big_map(n, 301)
    out = vt.VolMap(g, a).supersample().data.reshape(2 * n - 1, 2 * n - 1, 2 * n - 1)
    assert np.array_equal(bits(out[::2, ::2, ::2]), bits(a.reshape(n, n, n)))
    # interior x stencil (-y0 + 5 y1 + 5 y2 - y3)/8 evaluated exactly as the reference does, on one row
    src = a.reshape(n, n, n)[10, 20, :].astype(np.float32)
    y0, y1, y2, y3 = src[4], src[5], src[6], src[7]
    mu = np.float32(0.5); mu2 = np.float32(mu * mu)
    a0 = np.float32(np.float32(np.float32(y3 - y2) - y0) + y1)
    a1 = np.float32(np.float32(y0 - y1) - a0)
    a2 = np.float32(y2 - y0)
    want = np.float32(np.float32(np.float32(np.float32(np.float32(a0 * mu) * mu2) + np.float32(a1 * mu2)) + np.float32(a2 * mu)) + y1)
    assert out[20, 40, 11] == want
    del out
    col = np.random.default_rng(302).random(n * n, dtype=np.float32)
    flat = np.repeat(col, n)                          # constant along x
    o2 = vt.VolMap(g, flat).supersample().data.reshape(2 * n - 1, 2 * n - 1, 2 * n - 1)
    assert np.array_equal(bits(o2[:, :, 0:1].repeat(2 * n - 1, 2)), bits(o2))


def test_downsample_then_constant_and_checksum():
    n = 512
    g, a = big_map(n, 303)
    d = vt.VolMap(g, a).downsample().data.reshape(n // 2, n // 2, n // 2)
    # 2x2x2 box mean in double, in the reference's order, on a sample of voxels
    s = a.reshape(n, n, n)
    for (z, y, x) in [(0, 0, 0), (255, 255, 255), (17, 200, 93), (128, 1, 254)]:
        acc = 0.0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    acc += float(s[2 * z + dz, 2 * y + dy, 2 * x + dx])
        assert d[z, y, x] == np.float32(acc * 0.125)
    # a checksum of checksums: the mean of the downsampled map is the mean of the map (to rounding)
    assert abs(float(d.astype(np.float64).mean()) - float(a.astype(np.float64).mean())) < 1e-7


def test_streams_idempotence_and_cc_properties_512():
    n = 512
    g, a = big_map(n, 304)
    A = vt.VolMap(g, a)
    assert vt.cc_threaded(A, A) == pytest.approx(1.0, abs=1e-12)
    M = A.clone().binmask(0.5)
    once = M.data
    assert np.array_equal(bits(M.binmask(0.5).data), bits(once))            # idempotent
    assert np.array_equal(once != 0, a > 0.5)
    C = A.clone().clamp(0.25, 0.75)
    c1 = C.data
    assert np.array_equal(bits(C.clamp(0.25, 0.75).data), bits(c1))
    assert np.array_equal(bits(c1), bits(np.clip(a, np.float32(0.25), np.float32(0.75))))
    # CC is invariant under a positive affine map of B (up to rounding of the scaled values)
    b = (0.7 * a + 0.3 * np.random.default_rng(305).random(g.n, dtype=np.float32)).astype(np.float32)
    B = vt.VolMap(g, b)
    cc = vt.cc_threaded(A, B)
    B2 = B.clone().scale_by(2.0)                      # exact in float
    assert vt.cc_threaded(A, B2) == pytest.approx(cc, abs=1e-12)
    # add on one lattice is the float sum, bit for bit
    S = vt.add(A, B)
    assert np.array_equal(bits(S.data), bits(a + b))
    # the thresholded voxel count is the count of voxels of A at or above the threshold
    _, nvox, _ = vt.cc_threaded(A, B, 0.5, full=True)
    assert nvox == int(np.count_nonzero(a >= np.float32(0.5)))


def test_blur_large_properties():
    """512 x 256 x 128, sigma 2: a constant map stays constant up to the rounding of the tap sums, the
    blur of a map that is constant along z equals the 2-D blur broadcast along z, and the separable
    passes commute with transposition of the data (x <-> y) only up to rounding -- so compare one
    interior voxel with the tap sums evaluated in the documented order instead"""
    nx, ny, nz = 512, 256, 128
    g = vt.Grid.make((0, 0, 0), (nx, ny, nz), 1.0)
    a = np.random.default_rng(306).random(g.n, dtype=np.float32)
    out = vt.VolMap(g, a).gaussian_blur(2.0).data.reshape(nz, ny, nx)
    R = 6
    w = np.exp(-0.5 * (np.arange(-R, R + 1) / 2.0) ** 2)
    w = (w / w.sum()).astype(np.float32)
    src = a.reshape(nz, ny, nx)

    def fma(x, y, z):   # single rounding, like fmaf
        return np.float32(np.float64(x) * np.float64(y) + np.float64(z))

    def blur1(vals):
        acc = np.float32(0.0)
        for k in range(2 * R + 1):
            acc = fma(w[k], vals[k], acc)
        return acc

    z0, y0, x0 = 40, 100, 300
    xb = np.empty((2 * R + 1, 2 * R + 1), np.float32)
    for iz in range(2 * R + 1):
        for iy in range(2 * R + 1):
            xb[iz, iy] = blur1(src[z0 - R + iz, y0 - R + iy, x0 - R:x0 + R + 1])
    yb = np.array([blur1(xb[iz, :]) for iz in range(2 * R + 1)], np.float32)
    want = blur1(yb)
    # the taps themselves are computed by the library in float; allow their last-bit differences
    assert abs(float(out[z0, y0, x0]) - float(want)) <= 4e-7
    plane = np.random.default_rng(307).random(nx * ny, dtype=np.float32)
    o2 = vt.VolMap(g, np.tile(plane, nz)).gaussian_blur(2.0).data.reshape(nz, ny, nx)
    assert np.max(np.abs(o2 - o2[nz // 2][None])) <= 2e-7     # constant along z up to the z-tap rounding


def test_fit_overlap_equals_serial(monkeypatch):
    """the two-stream pose pipeline (list building of batch i+1 beside splat + cc of batch i) returns the
    same bits as the serial one, across several batches and a ragged tail"""
    pos, radii, mass = make_structure(6000, 24.0, seed=51, center=(40.0, 40.0, 40.0))
    radscale, gsp, quality = vt.sim_policy(5.0)
    com = vt.measure_center(pos, mass)
    T = vt.sim(vt.pose_apply(pos, com, (24.0, 48.0, 0.0)), radii, quality, radscale, 1.0)
    T.upload(add_noise(T.data, 0.02, seed=52))
    rng = np.random.default_rng(53)
    rots = (rng.integers(0, 15, size=(300, 3)) * 24).astype(np.float32)
    shifts = rng.integers(-1, 2, size=(300, 3)).astype(np.float32)
    poses = vt.FitContext.make_poses(rots, shifts)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("VOLTOOL_FIT_OVERLAP", mode)
        monkeypatch.setenv("VOLTOOL_FIT_BATCH", "64")
        ctx = vt.FitContext(pos, radii, T, 5.0)
        out[mode] = ctx.eval_poses(com, poses).copy()
        again = ctx.eval_poses(com, poses)
        assert np.array_equal(bits(again), bits(out[mode]))                 # deterministic
        ctx.close()
    assert np.array_equal(bits(out["0"]), bits(out["1"]))
    assert np.isfinite(out["1"]).all() and out["1"].max() > 0.5


def test_fit_list_overflow_falls_back_to_cooperative_splat(monkeypatch):
    """a tile list that does not fit its slot is never truncated: the call notices the flag and redoes the
    poses with the block-cooperative splat kernel -- the bits of a context that was told to use that kernel
    from the start, and the list kernel's values to rounding"""
    pos, radii, mass = make_structure(5000, 22.0, seed=61, center=(36.0, 36.0, 36.0))
    radscale, gsp, quality = vt.sim_policy(5.0)
    com = vt.measure_center(pos, mass)
    T = vt.sim(vt.pose_apply(pos, com, (48.0, 0.0, 24.0)), radii, quality, radscale, 1.0)
    rots = (np.random.default_rng(62).integers(0, 15, size=(150, 3)) * 24).astype(np.float32)
    poses = vt.FitContext.make_poses(rots)
    monkeypatch.setenv("VOLTOOL_FIT_BATCH", "64")
    ctx = vt.FitContext(pos, radii, T, 5.0)
    want = ctx.eval_poses(com, poses).copy()
    ctx.close()
    monkeypatch.setenv("VOLTOOL_LIST_CAP", "32")          # far below the fullest tile
    ctx = vt.FitContext(pos, radii, T, 5.0)
    got = ctx.eval_poses(com, poses).copy()
    again = ctx.eval_poses(com, poses)                       # the context stays on the fallback
    ctx.close()
    assert np.array_equal(bits(again), bits(got))
    # the two splat kernels group their partial sums differently: same support, values within the float
    # rounding of a few dozen additions (both are within 1e-5 of the oracle, test_gpu_parity)
    assert np.max(np.abs(got - want) / np.abs(want)) < 2e-6
    monkeypatch.delenv("VOLTOOL_LIST_CAP")
    monkeypatch.setenv("VOLTOOL_SPLAT_COOP", "1")
    ctx = vt.FitContext(pos, radii, T, 5.0)
    coop = ctx.eval_poses(com, poses)
    ctx.close()
    assert np.array_equal(bits(coop), bits(got))             # fallback == cooperative kernel from the start


def test_sim_list_estimate_and_overflow(monkeypatch):
    """vt_sim sizes its tile lists from the mean tile fill (no counting pass); a list that does not fit makes
    the call fall back to the cooperative kernel -- also for a structure far denser in one corner than on
    average, which is what defeats the estimate"""
    pos, radii, _ = make_structure(20000, 30.0, seed=71)
    clump = (np.random.default_rng(72).random((6000, 3)) * 6.0 + 70.0).astype(np.float32)   # 6000 atoms in a 6 A cube
    pos2 = np.concatenate([pos, clump]).astype(np.float32)
    rad2 = np.concatenate([radii, np.full(6000, 1.5, np.float32)])
    radscale, gsp, quality = vt.sim_policy(5.0)
    for p_, r_ in ((pos, radii), (pos2, rad2)):
        want = vt.sim(p_, r_, quality, radscale, gsp)
        monkeypatch.setenv("VOLTOOL_SPLAT_COOP", "1")
        coop = vt.sim(p_, r_, quality, radscale, gsp)
        monkeypatch.delenv("VOLTOOL_SPLAT_COOP")
        monkeypatch.setenv("VOLTOOL_LIST_CAP", "32")
        forced = vt.sim(p_, r_, quality, radscale, gsp)
        monkeypatch.delenv("VOLTOOL_LIST_CAP")
        a, b, c = want.data, coop.data, forced.data
        assert want.grid.astuple() == coop.grid.astuple() == forced.grid.astuple()
        assert np.array_equal(bits(b), bits(c))                       # forced overflow == cooperative kernel
        assert np.array_equal(a != 0, b != 0)                         # identical support
        nz = a != 0
        assert np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz])) < 1e-5      # each kernel is within 1e-5 of the oracle


def test_fit_rejects_skewed_target():
    pos, radii, mass = make_structure(300, 8.0, seed=81)
    g = vt.Grid.make((0.0, 0.0, 0.0), (24, 24, 24), axes=((23.0, 1.0, 0.0), (0.0, 23.0, 0.5), (0.0, 0.0, 23.0)))
    T = vt.VolMap(g, np.random.default_rng(82).random(g.n, dtype=np.float32))
    with pytest.raises(vt.VoltoolGpuError):
        vt.FitContext(pos, radii, T, 5.0)


# ------------------------------------------------------------------------------------------ round 2 kernels at size
def test_blur_one_pass_tma_equals_two_kernel_form_large(monkeypatch):
    """384 x 264 x 200 (ragged 16 x 64 tiles), sigma 0.6 / 1 / 2: the one-pass tensor-map TMA form (default up to
    sigma 1) and the two-kernel form produce the same bits"""
    nx, ny, nz = 384, 264, 200
    g = vt.Grid.make((0, 0, 0), (nx, ny, nz), 1.0)
    a = np.random.default_rng(401).random(g.n, dtype=np.float32)
    A = vt.VolMap(g, a)
    for sigma in (0.6, 1.0, 2.0):
        monkeypatch.setenv("VOLTOOL_BLUR_MODE", "tma")
        one = A.clone().gaussian_blur(sigma).data
        monkeypatch.setenv("VOLTOOL_BLUR_MODE", "two")
        two = A.clone().gaussian_blur(sigma).data
        monkeypatch.delenv("VOLTOOL_BLUR_MODE")
        dflt = A.clone().gaussian_blur(sigma).data
        assert np.array_equal(bits(one), bits(two)) and np.array_equal(bits(dflt), bits(two)), sigma


def test_lean_correlate_equals_ring_kernel_large(monkeypatch):
    """320^3 x 320^3, B offset by half a voxel in x: the packed regular-sampling kernel (tensor-map TMA ring) and the
    general ring kernel see the same samples -- identical voxel count, sums equal to the order of the float partials,
    and identical bits for the two-map ops"""
    n = 320
    g, a = big_map(n, 402)
    gb = vt.Grid.make((0.5, 0, 0), (n, n, n), 1.0)
    b = np.random.default_rng(403).random(gb.n, dtype=np.float32)
    A, B = vt.VolMap(g, a), vt.VolMap(gb, b)
    for thr in (-FLT_MAX, 0.5):
        cc1, n1, s1 = vt.cc_threaded(A, B, thr, full=True)
        monkeypatch.setenv("VOLTOOL_CC_NOLEAN", "1")
        cc0, n0, s0 = vt.cc_threaded(A, B, thr, full=True)
        monkeypatch.delenv("VOLTOOL_CC_NOLEAN")
        assert n1 == n0 and abs(cc1 - cc0) <= 1e-7
        assert np.allclose(s1, s0, rtol=1e-6)
    for op in (vt.add, vt.multiply):
        o1 = op(A, B).data
        monkeypatch.setenv("VOLTOOL_CC_NOLEAN", "1")
        o0 = op(A, B).data
        monkeypatch.delenv("VOLTOOL_CC_NOLEAN")
        assert np.array_equal(bits(o1), bits(o0))


def test_vol_com_bit_exact_on_a_noisy_padded_256_map():
    """the cfg2-style target (density in the middle of a zero-padded 256^3 map + signed noise everywhere): the float
    sums random-walk through the noise, which is the hard case for the parallel evaluation; bits must equal the
    sequential loop's (numpy float32 cumulative arithmetic in voxel order)"""
    n = 256
    g = vt.Grid.make((0, 0, 0), (n, n, n), 1.0)
    r = np.random.default_rng(404)
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    blob = np.exp(-((x - 130.0) ** 2 + (y - 120.0) ** 2 + (z - 125.0) ** 2) / (2 * 30.0 ** 2)).astype(np.float32)
    d = (blob + np.float32(0.02) * r.standard_normal((n, n, n)).astype(np.float32)).astype(np.float32).reshape(-1)
    com = vt.VolMap(g, d).vol_com()
    # sequential float32 sums in voxel order: np.add.accumulate keeps float32 and adds left to right
    want = np.empty(3, np.float32)
    mass = np.float64(0.0)
    for k, coord in enumerate((x, y, z)):
        addend = (d * coord.reshape(-1).astype(np.float32)).astype(np.float32)
        want[k] = np.add.accumulate(addend, dtype=np.float32)[-1]
    mass = np.add.accumulate(d.astype(np.float64))[-1]
    scale = np.float32(1.0 / mass)
    want = (scale * want).astype(np.float32)
    assert np.array_equal(bits(com), bits(want)), (com, want)
