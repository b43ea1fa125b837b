"""Generates tests/golden/voltool_golden.npz from the reference's OWN object code (oracle/_ref:
VolumetricData.C / Voltool.C / MDFF.C compiled unchanged) for everything that is in the snapshot,
and from the documented restatements (splat, blur, fit driver -- with the reference's cc_threaded
as CC backend) for what is not.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The .npz travels to the GPU box, where /root/reference does not exist."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref, Grid  # noqa: E402
from tests.synth import make_structure, make_target  # noqa: E402

FLT_MAX = 3.4028234663852886e38


def gpack(g):
    return np.array(list(g.origin) + list(g.xaxis) + list(g.yaxis) + list(g.zaxis) + [g.xsize, g.ysize, g.zsize], np.float64)


def main():
    ref = Ref()
    ref.set_numprocs(1)
    out = {}
    r = np.random.default_rng(20261019)
    grids = [Grid.make((0.0, 0.0, 0.0), (12, 10, 9), 1.0), Grid.make((0.5, -1.25, 2.0), (9, 13, 8), (1.1, 0.9, 1.3)),
             Grid.make((-1.3, 0.7, 0.4), (14, 7, 11), 0.77)]
    data = [(r.random(g.n, dtype=np.float32) * 2 - 0.5).astype(np.float32) for g in grids]
    data[1][::37] = np.nan
    for i, (g, d) in enumerate(zip(grids, data)):
        out[f"grid{i}"] = gpack(g)
        out[f"data{i}"] = d
    # cc_threaded (reference object code), both thresholds
    for i in range(3):
        for j in range(3):
            for t, thr in enumerate((-FLT_MAX, 0.25)):
                out[f"cc_{i}{j}_{t}"] = np.float64(ref.ref_cc(grids[i], data[i], grids[j], data[j], thr))
    # add / subtract / multiply / average (reference object code)
    for op in range(4):
        for interp in (0, 1):
            for uni in (0, 1):
                g, o = ref.ref_binary(op, grids[0], data[0], grids[2], data[2], bool(interp), bool(uni))
                out[f"bin_{op}{interp}{uni}_grid"] = gpack(g)
                out[f"bin_{op}{interp}{uni}_data"] = o
    # unary chain + cache, stats, com, histogram, resizing (reference object code)
    g = grids[0]
    v = ref.vol(g, data[0])
    chain = [("clamp", (-0.2, 1.2)), ("scale_by", (-1.5,)), ("scalar_add", (0.75,)), ("rescale_range", (0.0, 1.0)),
             ("sigma_scale", ()), ("mdff_potential", (0.2,)), ("binmask", (0.5,))]
    for k, (name, p) in enumerate(chain):
        v.op(name, *p)
        out[f"unary_{k}_{name}"] = v.data
        out[f"unary_{k}_cache"] = np.array(v.cache.astuple(), np.float64)
    v = ref.vol(g, data[0])
    out["stats"] = np.array([v.mean(), v.sigma(), *v.datarange()], np.float32)
    out["com"] = v.com()
    b, m = v.histogram(10)
    out["hist_bins"], out["hist_mid"] = b, m
    for name, p in (("pad", (2, -1, 0, 3, -2, 1)), ("crop", (1.2, 0.5, 0.9, 8.7, 12.0, 6.1)), ("downsample", ()), ("supersample", ())):
        v = ref.vol(g, data[0]).op(name, *p)
        out[f"{name}_grid"], out[f"{name}_data"] = gpack(v.grid), v.data
    # restatements (not in the snapshot): blur, splat, fit with the reference cc_threaded as backend
    out["blur15"] = ref.blur(g, data[0], 1.5)
    pos, radii, mass = make_structure(60, 5.0, seed=7, center=(10.0, 11.0, 12.0))
    out["atoms_pos"], out["atoms_radii"], out["atoms_mass"] = pos, radii, mass
    radscale, gsp, q = ref.sim_policy(6.0)
    sg, sd = ref.sim(pos, radii, q, radscale, gsp)
    out["sim_grid"], out["sim_data"] = gpack(sg), sd
    tg, td = make_target(ref, pos, radii, 6.0, 1.5, rot=(48, 24, 312), noise=0.01, seed=8, mass=mass, pad=3)
    out["target_grid"], out["target_data"] = gpack(tg), td
    tvol = ref.vol(tg, td)
    ref.install_cc_backend(tvol)
    out["calc_cc"] = np.float64(ref.calc_cc(pos, radii, tg, td, 6.0, 1))
    start = pos + np.array([1.0, -0.5, 0.75], np.float32)
    p_fit, res = ref.fit(start, radii, mass, tg, td, 6.0, direct=True, nthreads=1)
    ref.install_cc_backend(None)
    out["fit_start"], out["fit_pos"] = start, p_fit.reshape(-1, 3)
    out["fit_res"] = np.array([res.best_cc, res.n_evaluated, *res.stage_best], np.float64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "voltool_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
