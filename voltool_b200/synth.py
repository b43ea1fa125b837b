"""Synthetic workloads: protein-like atom balls and maps.  Pure numpy; no oracle."""
import numpy as np

# element mix: (fraction, radius A, mass)
ELEMENTS = [(0.50, 1.0, 1.008), (0.32, 1.5, 12.011), (0.08, 1.4, 14.007), (0.095, 1.3, 15.999), (0.005, 1.9, 32.06)]


def make_structure(natoms, ball_radius, seed, center=(50.0, 50.0, 50.0)):
    """natoms uniform in a ball; radii/masses by element category.  Returns pos (n,3), radii, mass (float32)"""
    r = np.random.default_rng(seed)
    v = r.normal(size=(natoms, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    rad = ball_radius * r.random(natoms) ** (1.0 / 3.0)
    pos = (np.asarray(center)[None, :] + v * rad[:, None]).astype(np.float32)
    u = r.random(natoms)
    edges = np.cumsum([e[0] for e in ELEMENTS])
    cat = np.minimum(np.searchsorted(edges, u), len(ELEMENTS) - 1)
    radii = np.array([ELEMENTS[c][1] for c in cat], np.float32)
    mass = np.array([ELEMENTS[c][2] for c in cat], np.float32)
    return pos, radii, mass


def add_noise(data, rel, seed):
    r = np.random.default_rng(seed)
    return (data + rel * float(data.max()) * r.standard_normal(data.shape).astype(np.float32)).astype(np.float32)
