"""The compiled reference-side adapter (shim/VolumetricDataGPU.C): the reference's OWN VolumetricData class, Voltool.C
free functions and cc_threaded, driven through the same flat API by two libraries --

    oracle/_ref/libvoltool_ref.so     the reference's object code, all CPU
    shim/_build/libvoltool_shim.so    the same object code with the hot-path bodies replaced by forwards to
                                      libvoltool_gpu.so (objcopy --weaken + strong definitions, shim/Makefile)

-- must leave every object in the same state: voxels, geometry and the cached-statistics block, bit for bit
(CC: only the order of the double sums differs).  Both libraries are built in the container that has /root/reference
and travel to the GPU box prebuilt."""
import os

import numpy as np
import pytest

import voltool_b200 as vt
from oracle.pyoracle import Grid, Ref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "shim", "_build", "libvoltool_shim.so")
FLT_MAX = 3.4028234663852886e38


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_grid(g1, g2):
    t1, t2 = g1.astuple(), g2.astuple()
    return t1[4] == t2[4] and all(np.array_equal(np.array(a), np.array(b), equal_nan=True) for a, b in zip(t1[:4], t2[:4]))


@pytest.fixture(scope="module")
def shim(ref):
    if not os.path.exists(SHIM):
        pytest.skip("shim/_build/libvoltool_shim.so not built (needs /root/reference: make -C shim)")
    vt.init(0)
    return Ref(SHIM)


def rng_map(seed, n):
    return np.random.default_rng(seed).random(n, dtype=np.float32)


def grids():
    return [Grid.make((0.0, 0.0, 0.0), (20, 18, 16), 1.0), Grid.make((0.5, -1.25, 2.0), (17, 21, 13), (1.1, 0.9, 1.3)),
            Grid.make((-3.3, 1.7, 0.4), (25, 12, 19), 0.77)]


def same_state(a, b):
    assert same_grid(a.grid, b.grid)
    assert np.array_equal(bits(a.data), bits(b.data))
    assert a.cache.astuple() == b.cache.astuple(), (a.cache.astuple(), b.cache.astuple())


CHAINS = [
    [("clamp", 0.2, 0.8), ("scale_by", -1.5), ("scalar_add", 0.25), ("rescale_range", 0.0, 1.0)],
    [("sigma_scale",), ("binmask", 0.1)],
    [("mdff_potential", 0.3), ("scale_by", 2.0)],
    [("pad", 2, 1, 0, 3, 1, 1), ("downsample",), ("supersample",)],
    [("crop", 2.0, 1.0, 3.0, 11.0, 9.5, 10.0), ("blur", 1.2), ("pad", -1, -1, -2, 0, -1, -1)],
    [("supersample",), ("blur", 0.7), ("downsample",), ("clamp", 0.3, 0.6)],
]


@pytest.mark.parametrize("chain", range(len(CHAINS)))
def test_volumetricdata_methods_same_state(ref, shim, chain):
    """every mutating VolumetricData method the adapter replaces, in chains (so stale caches carry over)"""
    for k, g in enumerate(grids()):
        d = rng_map(300 + k, g.n)
        a, b = ref.vol(g, d), shim.vol(g, d)
        same_state(a, b)
        for op in CHAINS[chain]:
            a.op(op[0], *op[1:]); b.op(op[0], *op[1:])
            same_state(a, b)
            # the statistics calls and what they leave in the cache
            assert a.datarange() == b.datarange()
            assert bits(np.float32(a.mean())) == bits(np.float32(b.mean()))
            assert abs(a.sigma() - b.sigma()) <= 2e-7 * abs(a.sigma())     # tree vs sequential double sum
        a.close(); b.close()


def test_cc_threaded_and_two_map_ops(ref, shim):
    gs = grids()
    for i, gA in enumerate(gs):
        for j, gB in enumerate(gs):
            a, b = rng_map(10 + i, gA.n), rng_map(20 + j, gB.n)
            for thr in (-FLT_MAX, 0.3):
                r, s = ref.ref_cc(gA, a, gB, b, thr), shim.ref_cc(gA, a, gB, b, thr)
                assert (np.isnan(r) and np.isnan(s)) or abs(r - s) <= 1e-9
            for op in range(4):
                for interp in (True, False):
                    for union in (False, True):
                        g1, o1 = ref.ref_binary(op, gA, a, gB, b, interp, union)
                        g2, o2 = shim.ref_binary(op, gA, a, gB, b, interp, union)
                        assert same_grid(g1, g2) and np.array_equal(bits(o1), bits(o2)), (op, interp, union)


def test_vol_com_and_histogram(ref, shim):
    for k, g in enumerate(grids()):
        d = rng_map(70 + k, g.n)
        a, b = ref.vol(g, d), shim.vol(g, d)
        assert np.array_equal(bits(a.com()), bits(b.com()))
        ha, ma = a.histogram(12)
        hb, mb = b.histogram(12)
        assert np.array_equal(ha, hb) and np.array_equal(bits(ma), bits(mb))
        assert a.cache.astuple() == b.cache.astuple()


def test_fit_driver_over_adapter_cc(oracle, ref, shim):
    """the oracle's restated fit() with its CC served by cc_threaded: the reference's object code vs the adapter's"""
    from tests.synth import make_structure, make_target
    pos, radii, mass = make_structure(300, 9.0, seed=91)
    g, t = make_target(oracle, pos, radii, 6.0, 1.5, rot=(24, 48, 0), noise=0.02, seed=92, mass=mass)
    out = []
    for lib in (ref, shim):
        tv = lib.vol(g, t)
        lib.install_cc_backend(tv)
        try:
            out.append(lib.fit(pos, radii, mass, g, t, 6.0, direct=False, nthreads=1))
        finally:
            lib.install_cc_backend(None)
    (p1, r1), (p2, r2) = out
    assert list(r1.stage_best) == list(r2.stage_best) and r1.n_evaluated == r2.n_evaluated
    assert np.array_equal(bits(p1), bits(p2)) and abs(r1.best_cc - r2.best_cc) <= 1e-6
