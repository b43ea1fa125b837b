"""Golden vectors generated from the reference's own object code (tests/golden/make_golden.py).
CPU half: the oracle restatement reproduces them (works where oracle/_ref cannot be rebuilt).
GPU half: the CUDA path, through the C ABI, reproduces them."""
import os

import numpy as np
import pytest

import voltool_b200 as vt
from oracle.pyoracle import Grid as OGrid

HERE = os.path.dirname(os.path.abspath(__file__))
FLT_MAX = 3.4028234663852886e38
CHAIN = [("clamp", (-0.2, 1.2)), ("scale_by", (-1.5,)), ("scalar_add", (0.75,)), ("rescale_range", (0.0, 1.0)),
         ("sigma_scale", ()), ("mdff_potential", (0.2,)), ("binmask", (0.5,))]


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(HERE, "golden", "voltool_golden.npz")))


def ungrid(a):
    g = OGrid()
    for k in range(3):
        g.origin[k], g.xaxis[k], g.yaxis[k], g.zaxis[k] = a[k], a[3 + k], a[6 + k], a[9 + k]
    g.xsize, g.ysize, g.zsize = int(a[12]), int(a[13]), int(a[14])
    return g


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_grid(g, packed):
    t = g.astuple()
    flat = np.array(list(t[0]) + list(t[1]) + list(t[2]) + list(t[3]) + list(t[4]), np.float64)
    return np.array_equal(flat, packed, equal_nan=True)


# ------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_reproduces_golden(oracle, gold):
    grids = [ungrid(gold[f"grid{i}"]) for i in range(3)]
    data = [gold[f"data{i}"] for i in range(3)]
    for i in range(3):
        for j in range(3):
            for t, thr in enumerate((-FLT_MAX, 0.25)):
                want = float(gold[f"cc_{i}{j}_{t}"])
                got = oracle.cc(grids[i], data[i], grids[j], data[j], thr)
                assert (np.isnan(want) and np.isnan(got)) or got == want
    for op in range(4):
        for interp in (0, 1):
            for uni in (0, 1):
                g, o = oracle.binary(op, grids[0], data[0], grids[2], data[2], bool(interp), bool(uni))
                assert same_grid(g, gold[f"bin_{op}{interp}{uni}_grid"])
                assert np.array_equal(bits(o), bits(gold[f"bin_{op}{interp}{uni}_data"]))
    g = grids[0]
    d = data[0].copy()
    c = oracle.new_cache(g, d)
    for k, (name, p) in enumerate(CHAIN):
        d = oracle.unary(name, g, d, c, *p)
        assert np.array_equal(bits(d), bits(gold[f"unary_{k}_{name}"]))
        assert np.array_equal(np.array(c.astuple(), np.float64), gold[f"unary_{k}_cache"], equal_nan=True)
    c = oracle.new_cache(g, data[0])
    st = np.array([oracle.mean(g, data[0], c), oracle.sigma(g, data[0], c), *oracle.datarange(g, data[0], c)], np.float32)
    assert np.array_equal(bits(st), bits(gold["stats"]))
    assert np.array_equal(bits(oracle.vol_com(g, data[0])), bits(gold["com"]))
    b, m = oracle.histogram(g, data[0], oracle.new_cache(g, data[0]), 10)
    assert np.array_equal(b, gold["hist_bins"]) and np.array_equal(bits(m), bits(gold["hist_mid"]))
    for name, fn, p in (("pad", oracle.pad, ((2, -1, 0, 3, -2, 1),)), ("crop", oracle.crop, ((1.2, 0.5, 0.9, 8.7, 12.0, 6.1),)),
                        ("downsample", oracle.downsample, ()), ("supersample", oracle.supersample, ())):
        g2, o = fn(g, data[0], *p)
        assert same_grid(g2, gold[f"{name}_grid"]) and np.array_equal(bits(o), bits(gold[f"{name}_data"]))
    assert np.array_equal(bits(oracle.blur(g, data[0], 1.5)), bits(gold["blur15"]))
    pos, radii, mass = gold["atoms_pos"], gold["atoms_radii"], gold["atoms_mass"]
    radscale, gsp, q = oracle.sim_policy(6.0)
    sg, sd = oracle.sim(pos, radii, q, radscale, gsp)
    assert same_grid(sg, gold["sim_grid"]) and np.array_equal(bits(sd), bits(gold["sim_data"]))
    tg, td = ungrid(gold["target_grid"]), gold["target_data"]
    # the port's own cc here (the golden used the reference's cc_threaded): same sequential sums
    assert oracle.calc_cc(pos, radii, tg, td, 6.0, 1) == float(gold["calc_cc"])
    p_fit, res = oracle.fit(gold["fit_start"], radii, mass, tg, td, 6.0, direct=True, nthreads=1)
    assert np.array_equal(bits(p_fit.reshape(-1, 3)), bits(gold["fit_pos"]))
    assert [res.best_cc, res.n_evaluated, *res.stage_best] == [np.float32(gold["fit_res"][0])] + [int(x) for x in gold["fit_res"][1:]]


# ------------------------------------------------------------------------------------------ GPU: product
@pytest.mark.gpu
def test_cuda_reproduces_golden(gold):
    vt.init(0)
    G = vt.Grid.from_fields
    grids = [ungrid(gold[f"grid{i}"]) for i in range(3)]
    data = [gold[f"data{i}"] for i in range(3)]
    maps = [vt.VolMap(G(g), d) for g, d in zip(grids, data)]
    for i in range(3):
        for j in range(3):
            for t, thr in enumerate((-FLT_MAX, 0.25)):
                want = float(gold[f"cc_{i}{j}_{t}"])
                got = vt.cc_threaded(maps[i], maps[j], thr)
                assert (np.isnan(want) and np.isnan(got)) or abs(got - want) <= 1e-9, (i, j, t)
    for op in range(4):
        for interp in (0, 1):
            for uni in (0, 1):
                o = vt.binary(op, maps[0], maps[2], bool(interp), bool(uni))
                assert same_grid(o.grid, gold[f"bin_{op}{interp}{uni}_grid"])
                assert np.array_equal(bits(o.data), bits(gold[f"bin_{op}{interp}{uni}_data"])), (op, interp, uni)
    m = vt.VolMap(G(grids[0]), data[0])
    names = {"rescale_range": "rescale_voxel_value_range"}
    for k, (name, p) in enumerate(CHAIN):
        getattr(m, names.get(name, name))(*p)
        assert np.array_equal(bits(m.data), bits(gold[f"unary_{k}_{name}"])), name
        assert np.array_equal(np.array(m.cache.astuple(), np.float64), gold[f"unary_{k}_cache"], equal_nan=True), name
    m = vt.VolMap(G(grids[0]), data[0])
    st = np.array([m.mean(), m.sigma(), *m.datarange()], np.float32)
    assert np.allclose(st, gold["stats"], rtol=1.2e-7, atol=0)
    assert np.array_equal(bits(m.vol_com()), bits(gold["com"]))
    b, mid = m.histogram(10)
    assert np.array_equal(b, gold["hist_bins"]) and np.array_equal(bits(mid), bits(gold["hist_mid"]))
    for name, call in (("pad", lambda v: v.pad(2, -1, 0, 3, -2, 1)), ("crop", lambda v: v.crop(1.2, 0.5, 0.9, 8.7, 12.0, 6.1)),
                       ("downsample", lambda v: v.downsample()), ("supersample", lambda v: v.supersample())):
        v = call(vt.VolMap(G(grids[0]), data[0]))
        assert same_grid(v.grid, gold[f"{name}_grid"]) and np.array_equal(bits(v.data), bits(gold[f"{name}_data"])), name
    assert np.array_equal(bits(vt.VolMap(G(grids[0]), data[0]).gaussian_blur(1.5).data), bits(gold["blur15"]))
    pos, radii, mass = gold["atoms_pos"], gold["atoms_radii"], gold["atoms_mass"]
    radscale, gsp, q = vt.sim_policy(6.0)
    s = vt.sim(pos, radii, q, radscale, gsp)
    assert same_grid(s.grid, gold["sim_grid"])
    want = gold["sim_data"]
    assert np.array_equal(s.data == 0, want == 0)
    assert (np.abs(s.data - want) / np.maximum(np.abs(want), 1e-30)).max() < 1e-5
    target = vt.VolMap(G(ungrid(gold["target_grid"])), gold["target_data"])
    assert abs(vt.calc_cc(pos, radii, target, 6.0) - float(gold["calc_cc"])) <= 1e-5 * abs(float(gold["calc_cc"]))
    p_fit, res = vt.fit(gold["fit_start"], radii, mass, target, 6.0)
    assert np.array_equal(bits(p_fit), bits(gold["fit_pos"]))
    assert [res.n_evaluated, *res.stage_best] == [int(x) for x in gold["fit_res"][1:]]
    assert abs(res.best_cc - gold["fit_res"][0]) <= 1e-5 * abs(gold["fit_res"][0])
