"""Test-side synthetic inputs; targets are generated with the CPU oracle (tests only)."""
import numpy as np

from voltool_b200.synth import make_structure, add_noise  # noqa: F401


def make_target(orc, pos, radii, resolution=5.0, spacing=1.0, rot=(24, 48, 0), noise=0.02, seed=2, pad=3,
                mass=None):
    """density of the structure rotated by `rot` degrees (X,Y,Z about its centre), sampled at `spacing`,
    zero-padded by `pad` voxels, plus Gaussian noise.  Returns (oracle Grid, float32 data)."""
    w = np.ones(len(radii), np.float32) if mass is None else mass
    com = orc.measure_center(pos, w)
    p = orc.pose_transform(pos, com, 1, int(rot[0]), int(rot[1]), int(rot[2]))
    radscale, _, quality = orc.sim_policy(resolution)
    g, d = orc.sim(p, radii, quality, radscale, spacing, nthreads=4)
    if pad:
        g, d = orc.pad(g, d, (pad,) * 6)
    if noise:
        d = add_noise(d, noise, seed)
    return g, d
