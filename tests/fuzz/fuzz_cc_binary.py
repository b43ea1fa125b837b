"""GPU fuzzer (not collected by pytest): random pairs of orthogonal grids (sizes, spacings, offsets); the CC of
the pair against the oracle (1e-9, same voxel count) on the direct, ring and identity paths, and the four
two-map ops in all interp x union modes bit for bit.  SEED=.. N=.. python tests/fuzz/fuzz_cc_binary.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import voltool_b200 as vt
from oracle.pyoracle import Oracle, Grid as OGrid
vt.init(0)
orc = Oracle()
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
N = int(os.environ.get("N", "40"))
FLT_MAX = 3.4028234663852886e38
def bits(a): return np.ascontiguousarray(a, np.float32).view(np.uint32)
bad = 0
for it in range(N):
    sa = tuple(int(x) for x in rng.integers(4, 70, size=3))
    sb = tuple(int(x) for x in rng.integers(4, 70, size=3))
    spa = float(rng.choice([0.5, 0.77, 1.0, 1.0, 1.3, 1.5]))
    spb = float(rng.choice([0.5, 1.0, 1.0, 1.06, 1.5, 2.0]))
    oa = tuple(rng.uniform(-3, 3, size=3))
    ob = tuple(rng.uniform(-3, 3, size=3)) if rng.random() < 0.7 else oa
    if rng.random() < 0.2: sb, spb, ob = sa, spa, oa            # one lattice
    A, B = OGrid.make(oa, sa, spa), OGrid.make(ob, sb, spb)
    a = (rng.random(A.n, dtype=np.float32) - 0.2).astype(np.float32)
    b = (rng.random(B.n, dtype=np.float32) - 0.2).astype(np.float32)
    if orc.intersection(A, B).n <= 0: continue
    GA, GB = vt.Grid.from_fields(A), vt.Grid.from_fields(B)
    for thr in (-FLT_MAX, 0.3):
        want, nw, _ = orc.cc(A, a, B, b, thr, nthreads=4, full=True)
        for env in ({}, {"VOLTOOL_CC_STAGED": "1"}, {"VOLTOOL_CC_IDENTITY": "1"}, {"VOLTOOL_FORCE_GENERIC": "1"}):
            for k in ("VOLTOOL_CC_STAGED", "VOLTOOL_CC_IDENTITY", "VOLTOOL_FORCE_GENERIC"): os.environ.pop(k, None)
            os.environ.update(env)
            got, ng, _ = vt.cc_threaded(vt.VolMap(GA, a), vt.VolMap(GB, b), thr, full=True)
            # the ring and identity kernels add the five sums in float over 4 rows before the double sum
            tol = 2e-7 if env and "GENERIC" not in next(iter(env)) else 1e-9
            same = (np.isnan(want) and np.isnan(got)) or abs(got - want) <= tol
            if ng != nw or not same:
                bad += 1; print("CC MISMATCH", sa, sb, spa, spb, thr, env, got, want, ng, nw, flush=True)
    for k in ("VOLTOOL_CC_STAGED", "VOLTOOL_CC_IDENTITY", "VOLTOOL_FORCE_GENERIC"): os.environ.pop(k, None)
    for op in range(4):
        for interp in (True, False):
            for use_union in (True, False):
                og, wantm = orc.binary(op, A, a, B, b, interp, use_union)
                if og.n <= 0 or og.n > 3_000_000: continue
                gotm = vt.binary(op, vt.VolMap(GA, a), vt.VolMap(GB, b), interp, use_union)
                if gotm.grid.astuple() != og.astuple() or not np.array_equal(bits(gotm.data), bits(wantm)):
                    bad += 1; print("BINARY MISMATCH", sa, sb, spa, spb, op, interp, use_union, flush=True)
print("fuzz cc/binary done:", N, "cases,", bad, "mismatches")
