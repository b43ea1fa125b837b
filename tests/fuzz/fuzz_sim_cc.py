"""GPU fuzzer (not collected by pytest): random structures, resolutions and spacings; vt_sim against the oracle
(identical support, 1e-5 relative) and batched pose CCs against the oracle's calc_cc (1e-5).
SEED=.. N=.. python tests/fuzz/fuzz_sim_cc.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import voltool_b200 as vt
from oracle.pyoracle import Oracle, Grid as OGrid
from voltool_b200.synth import make_structure
vt.init(0)
orc = Oracle()
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
N = int(os.environ.get("N", "30"))
bad = 0
worst = 0.0
for it in range(N):
    natoms = int(rng.integers(1, 6000))
    ball = float(rng.uniform(3.0, 28.0))
    pos, radii, mass = make_structure(natoms, ball, seed=int(rng.integers(1 << 30)), center=tuple(rng.uniform(-20, 60, size=3)))
    res = float(rng.choice([3.0, 4.0, 5.0, 8.0, 9.0, 12.0]))
    sp = float(rng.choice([0.0, 0.0, 1.0, 0.7, 1.9]))
    radscale, gsp, quality = orc.sim_policy(res, sp)
    g, want = orc.sim(pos, radii, quality, radscale, gsp, nthreads=8)
    if g.n > 6_000_000: continue
    m = vt.sim(pos, radii, quality, radscale, gsp)
    got = m.data
    ok = m.grid.astuple() == g.astuple() and np.array_equal(got == 0, want == 0)
    nz = want != 0
    err = float(np.max(np.abs(got[nz] - want[nz]) / np.abs(want[nz]))) if nz.any() else 0.0
    worst = max(worst, err)
    if not ok or err >= 1e-5:
        bad += 1; print("SIM MISMATCH", natoms, ball, res, sp, ok, err, flush=True)
    # pose CC of a few rigid poses against a noisy target of the same structure
    if natoms >= 50 and it % 3 == 0:
        com = orc.measure_center(pos, mass)
        t = (want + 0.02 * want.max() * rng.standard_normal(want.shape).astype(np.float32)).astype(np.float32)
        T = vt.VolMap(vt.Grid.from_fields(g), t)
        ctx = vt.FitContext(pos, radii, T, res)
        rots = (rng.integers(0, 15, size=(5, 3)) * 24).astype(np.float32)
        cc = ctx.eval_poses(com, vt.FitContext.make_poses(rots))
        for r, c in zip(rots, cc):
            p = orc.pose_transform(pos, com, 24, int(r[0]) // 24, int(r[1]) // 24, int(r[2]) // 24)
            w = np.float32(orc.calc_cc(p, radii, g, t, res, nthreads=8))
            if not (abs(float(c) - float(w)) <= 1e-5 * max(abs(float(w)), 1e-3) or (np.isnan(c) and np.isnan(w))):
                bad += 1; print("POSE CC MISMATCH", natoms, res, r, c, w, flush=True)
        ctx.close()
print("fuzz2 done:", N, "cases,", bad, "mismatches, worst sim rel err", worst)
