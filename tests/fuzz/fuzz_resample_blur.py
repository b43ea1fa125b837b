"""GPU fuzzer (not collected by pytest): random map sizes and cells; supersample (all three forms), Gaussian blur
(all three forms) and downsample against the CPU oracle, bit for bit.  SEED=.. N=.. python tests/fuzz/fuzz_resample_blur.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import voltool_b200 as vt
from oracle.pyoracle import Oracle, Grid as OGrid
vt.init(0)
orc = Oracle()
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
def bits(a): return np.ascontiguousarray(a, np.float32).view(np.uint32)
bad = 0
N = int(os.environ.get("N", "60"))
for it in range(N):
    size = tuple(int(x) for x in rng.integers(3, 150, size=3))
    if rng.random() < 0.3: size = (int(rng.integers(1, 40)) * 4, size[1], size[2])
    if size[0] * size[1] * size[2] > 1_500_000: size = (size[0], size[1], max(3, 1_500_000 // (size[0] * size[1])))
    g = OGrid.make(tuple(rng.normal(size=3)), size, tuple(rng.uniform(0.5, 2.0, size=3)))
    a = (rng.random(g.n, dtype=np.float32) * 2 - 0.3).astype(np.float32)
    G = vt.Grid.from_fields(g)
    # supersample
    og, want = orc.supersample(g, a)
    for mode in ("0", "1", "2"):
        os.environ["VOLTOOL_SUPERSAMPLE_FUSED"] = mode
        got = vt.VolMap(G, a).supersample().data
        if not np.array_equal(bits(got), bits(want)):
            bad += 1; print("SUPERSAMPLE MISMATCH", size, mode, flush=True)
    os.environ.pop("VOLTOOL_SUPERSAMPLE_FUSED")
    # blur
    sigma = float(rng.choice([0.4, 0.7, 1.0, 1.3, 1.7, 2.0]))
    wantb = orc.blur(g, a, sigma)
    for mode in ("", "yz", "xyz"):
        if mode: os.environ["VOLTOOL_BLUR_MODE"] = mode
        else: os.environ.pop("VOLTOOL_BLUR_MODE", None)
        got = vt.VolMap(G, a).gaussian_blur(sigma).data
        if not np.array_equal(bits(got), bits(wantb)):
            bad += 1; print("BLUR MISMATCH", size, sigma, repr(mode), flush=True)
    os.environ.pop("VOLTOOL_BLUR_MODE", None)
    # downsample
    dg, wantd = orc.downsample(g, a)
    if min(size) >= 2:
        got = vt.VolMap(G, a).downsample().data
        if not np.array_equal(bits(got), bits(wantd)):
            bad += 1; print("DOWNSAMPLE MISMATCH", size, flush=True)
print("fuzz done:", N, "cases,", bad, "mismatches")
