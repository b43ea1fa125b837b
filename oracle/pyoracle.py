"""ctypes bindings for the CPU checker.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under voltool_b200/ does.

  Oracle()  -> oracle/liboracle.so          (C restatement, voltool_oracle.c)
  Ref()     -> oracle/_ref/libvoltool_ref.so (the reference's own object code + glue); also
               exports every vo_* symbol, so Ref() is a superset of Oracle().
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)


class Grid(C.Structure):
    """mirrors vo_grid / VolumetricData.h:29-33"""
    _fields_ = [("origin", C.c_double * 3), ("xaxis", C.c_double * 3), ("yaxis", C.c_double * 3),
                ("zaxis", C.c_double * 3), ("xsize", C.c_int), ("ysize", C.c_int), ("zsize", C.c_int)]

    @classmethod
    def make(cls, origin, size, spacing=None, axes=None):
        g = cls()
        nx, ny, nz = [int(s) for s in size]
        g.xsize, g.ysize, g.zsize = nx, ny, nz
        for k in range(3):
            g.origin[k] = float(origin[k])
            g.xaxis[k] = g.yaxis[k] = g.zaxis[k] = 0.0
        if axes is not None:
            for k in range(3):
                g.xaxis[k], g.yaxis[k], g.zaxis[k] = float(axes[0][k]), float(axes[1][k]), float(axes[2][k])
        else:
            sp = (spacing,) * 3 if np.isscalar(spacing) else spacing
            g.xaxis[0] = float(sp[0]) * (nx - 1)
            g.yaxis[1] = float(sp[1]) * (ny - 1)
            g.zaxis[2] = float(sp[2]) * (nz - 1)
        return g

    @property
    def shape(self):
        return (self.zsize, self.ysize, self.xsize)

    @property
    def n(self):
        return self.xsize * self.ysize * self.zsize

    def copy(self):
        g = Grid()
        C.memmove(C.byref(g), C.byref(self), C.sizeof(Grid))
        return g

    def astuple(self):
        return (tuple(self.origin), tuple(self.xaxis), tuple(self.yaxis), tuple(self.zaxis),
                (self.xsize, self.ysize, self.zsize))


class Cache(C.Structure):
    _fields_ = [("minmax_valid", C.c_int), ("mean_valid", C.c_int), ("sigma_valid", C.c_int),
                ("cmin", C.c_float), ("cmax", C.c_float), ("cmean", C.c_float), ("csigma", C.c_float)]

    def astuple(self):
        return (self.minmax_valid, self.mean_valid, self.sigma_valid, self.cmin, self.cmax, self.cmean, self.csigma)


class FitResult(C.Structure):
    _fields_ = [("best_cc", C.c_float), ("com", C.c_float * 3), ("dencom", C.c_float * 3),
                ("n_evaluated", C.c_int), ("stage_best", C.c_int * 5)]


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def fp(a):
    return a.ctypes.data_as(c_float_p)


def build(ref=True):
    """(re)build the checker libraries with oracle/Makefile (gcc; no GPU involved)."""
    subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libvoltool_ref.so"))


class Oracle:
    LIB = os.path.join(HERE, "liboracle.so")

    def __init__(self, path=None):
        path = path or self.LIB
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.vo_cc.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Grid), c_float_p, C.c_double, c_double_p,
                            C.POINTER(C.c_long), c_double_p, C.c_int]
        L.vo_binary.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(Grid), c_float_p, C.POINTER(Grid), c_float_p,
                                C.POINTER(Grid), C.POINTER(c_float_p)]
        L.vo_free.argtypes = [C.c_void_p]
        for name in ("vo_clamp", "vo_rescale_range"):
            getattr(L, name).argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Cache), C.c_float, C.c_float]
        for name in ("vo_scale_by", "vo_scalar_add", "vo_binmask"):
            getattr(L, name).argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Cache), C.c_float]
        L.vo_sigma_scale.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Cache)]
        L.vo_mdff_potential.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Cache), C.c_double]
        L.vo_datarange.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Cache), c_float_p, c_float_p]
        L.vo_mean.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Cache)]
        L.vo_mean.restype = C.c_float
        L.vo_sigma.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Cache)]
        L.vo_sigma.restype = C.c_float
        L.vo_vol_com.argtypes = [C.POINTER(Grid), c_float_p, c_float_p]
        L.vo_vol_moveto.argtypes = [C.POINTER(Grid), c_float_p, c_float_p]
        L.vo_vol_move.argtypes = [C.POINTER(Grid), c_float_p]
        L.vo_histogram.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Cache), C.c_int, C.POINTER(C.c_long), c_float_p]
        L.vo_pad.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(c_float_p)] + [C.c_int] * 6
        L.vo_crop.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(c_float_p)] + [C.c_double] * 6
        L.vo_downsample.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(c_float_p)]
        L.vo_supersample.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(c_float_p)]
        L.vo_gaussian_blur.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(c_float_p), C.c_double]
        L.vo_blur_kernel.argtypes = [C.c_double, c_float_p, C.c_int]
        L.vo_sim.argtypes = [c_float_p, c_float_p, C.c_long, C.c_int, C.c_float, C.c_float, C.POINTER(Grid),
                             C.POINTER(c_float_p), C.c_int]
        L.vo_sim_policy.argtypes = [C.c_double, C.c_double, c_float_p, c_double_p, C.POINTER(C.c_int)]
        L.vo_measure_center.argtypes = [c_float_p, c_float_p, C.c_long, c_float_p]
        L.vo_pose_transform.argtypes = [c_float_p, C.c_long, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vo_pose_general.argtypes = [c_float_p, C.c_long, c_float_p, c_float_p, c_float_p]
        L.vo_calc_cc.argtypes = [c_float_p, c_float_p, C.c_long, C.POINTER(Grid), c_float_p, C.c_float, C.c_int]
        L.vo_calc_cc.restype = C.c_double
        L.vo_rotate_table.argtypes = [c_float_p, c_float_p, C.c_long, c_float_p, C.c_int, C.c_int, C.POINTER(Grid),
                                      c_float_p, C.c_float, c_float_p, C.c_int, C.c_int, C.c_int]
        L.vo_rotate_replay.argtypes = [c_float_p, C.c_int, c_float_p, C.POINTER(C.c_int)]
        L.vo_rotate_replay.restype = C.c_int
        L.vo_fit.argtypes = [c_float_p, c_float_p, c_float_p, C.c_long, C.POINTER(Grid), c_float_p, C.c_float,
                             C.POINTER(FitResult), C.c_int, C.c_int]
        L.vo_mat4_inverse.argtypes = [c_float_p]
        L.vo_mat4_multpoint3d.argtypes = [c_float_p, c_float_p, c_float_p]
        L.vo_grid_inverse.argtypes = [C.POINTER(Grid), C.c_int, c_float_p]
        L.vo_voxel_coord.argtypes = [C.POINTER(Grid), C.c_long, c_float_p, c_float_p, c_float_p]
        L.vo_voxel_coord_int.argtypes = [C.POINTER(Grid), C.c_int, c_float_p, c_float_p, c_float_p]
        L.vo_voxel_coord_from_cartesian.argtypes = [C.POINTER(Grid), c_float_p, c_float_p, C.c_int]
        L.vo_voxel_index_from_coord.argtypes = [C.POINTER(Grid), C.c_float, C.c_float, C.c_float]
        L.vo_voxel_index_from_coord.restype = C.c_long
        L.vo_value_from_coord.argtypes = [C.POINTER(Grid), c_float_p, C.c_float, C.c_float, C.c_float, C.c_int]
        L.vo_value_from_coord.restype = C.c_float
        L.vo_value_interp_from_coord.argtypes = [C.POINTER(Grid), c_float_p, C.c_float, C.c_float, C.c_float, C.c_int]
        L.vo_value_interp_from_coord.restype = C.c_float
        L.vo_init_from_intersection.argtypes = [C.POINTER(Grid)] * 3
        L.vo_init_from_union.argtypes = [C.POINTER(Grid)] * 3
        L.vo_cc_maps.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Grid), c_float_p, C.c_double, c_double_p,
                                 C.POINTER(Grid), C.POINTER(c_float_p), C.POINTER(Grid), C.POINTER(c_float_p)]
        L.vo_density_rotate.argtypes = [C.c_int, C.c_int, c_float_p, c_float_p, C.POINTER(C.c_int), C.POINTER(Grid),
                                        c_float_p, C.POINTER(Grid), c_float_p, c_float_p, C.c_int]
        L.vo_density_rotate.restype = C.c_int
        L.vo_density_reset.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(Grid), c_float_p, c_float_p, c_float_p]

    # ---- helpers -------------------------------------------------------------------------
    def _take(self, ptr, n):
        out = np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].copy()
        self.lib.vo_free(ptr)
        return out

    # ---- cc extras / map-rotating search (restatements, see voltool_oracle.c) ------------------
    def cc_maps(self, gA, a, gB, b, thr=-3.4028234663852886e38):
        """(cc, diff grid, diff, spatial grid, spatial): TclMDFF.C:180-184, 555-594"""
        a, b = f32(a), f32(b)
        cc, dg, sg = C.c_double(), Grid(), Grid()
        pd, ps = c_float_p(), c_float_p()
        self.lib.vo_cc_maps(C.byref(gA), fp(a), C.byref(gB), fp(b), thr, C.byref(cc), C.byref(dg), C.byref(pd),
                            C.byref(sg), C.byref(ps))
        return cc.value, dg, self._take(pd, dg.n), sg, self._take(ps, sg.n)

    def density_rotate(self, gA, a, gS, s, stride, max_rot, com, returncc=-1.0, nthreads=1):
        """density_rotate, TclVoltool.C:202-279: (returncc, bestrot or None, table)"""
        a, s, com = f32(a), f32(s), f32(com)
        rc = C.c_float(returncc)
        best = (C.c_int * 3)(-1, -1, -1)
        tab = np.full(max_rot ** 3, np.nan, np.float32)
        g2 = gS.copy()
        self.lib.vo_density_rotate(stride, max_rot, fp(com), C.byref(rc), best, C.byref(gA), fp(a), C.byref(g2), fp(s),
                                   fp(tab), nthreads)
        return rc.value, (None if best[0] < 0 else tuple(best)), tab

    def density_reset(self, gS, s, stride, bestrot, synthcom, com):
        s, synthcom, com = f32(s), f32(synthcom), f32(com)
        best = (C.c_int * 3)(*[int(v) for v in bestrot])
        g2 = gS.copy()
        self.lib.vo_density_reset(stride, best, C.byref(g2), fp(s), fp(synthcom), fp(com))
        return g2

    # ---- CC --------------------------------------------------------------------------------
    def cc(self, gA, a, gB, b, thr=-3.4028234663852886e38, nthreads=1, full=False):
        a, b = f32(a), f32(b)
        cc, n, sums = C.c_double(), C.c_long(), (C.c_double * 5)()
        self.lib.vo_cc(C.byref(gA), fp(a), C.byref(gB), fp(b), thr, C.byref(cc), C.byref(n), sums, nthreads)
        if full:
            return cc.value, n.value, list(sums)
        return cc.value

    def binary(self, op, gA, a, gB, b, interp=True, use_union=False):
        a, b = f32(a), f32(b)
        og, out = Grid(), c_float_p()
        self.lib.vo_binary(op, int(interp), int(use_union), C.byref(gA), fp(a), C.byref(gB), fp(b), C.byref(og),
                           C.byref(out))
        return og, self._take(out, og.n)

    # ---- unary (in place on a copy; returns data, cache) -------------------------------------
    def new_cache(self, g, d):
        """cache state right after the VolumetricData constructor (VolumetricData.C:52-63)"""
        c = Cache()
        mn, mx = C.c_float(), C.c_float()
        c.mean_valid = 1  # force the minmax-only branch, then restore
        self.lib.vo_datarange(C.byref(g), fp(d), C.byref(c), C.byref(mn), C.byref(mx))
        c.mean_valid = 0
        c.cmean = 0.0
        return c

    def unary(self, name, g, d, cache, *params):
        d = f32(d).copy()
        fn = getattr(self.lib, "vo_" + name)
        fn(C.byref(g), fp(d), C.byref(cache), *params)
        return d

    def datarange(self, g, d, cache):
        d = f32(d)
        mn, mx = C.c_float(), C.c_float()
        self.lib.vo_datarange(C.byref(g), fp(d), C.byref(cache), C.byref(mn), C.byref(mx))
        return mn.value, mx.value

    def mean(self, g, d, cache):
        return self.lib.vo_mean(C.byref(g), fp(f32(d)), C.byref(cache))

    def sigma(self, g, d, cache):
        return self.lib.vo_sigma(C.byref(g), fp(f32(d)), C.byref(cache))

    def vol_com(self, g, d):
        d = f32(d)
        com = np.zeros(3, np.float32)
        self.lib.vo_vol_com(C.byref(g), fp(d), fp(com))
        return com

    def vol_moveto(self, g, com, pos):
        g2, com, pos = g.copy(), f32(com), f32(pos)
        self.lib.vo_vol_moveto(C.byref(g2), fp(com), fp(pos))
        return g2

    def vol_move(self, g, mat):
        g2, mat = g.copy(), f32(mat)
        self.lib.vo_vol_move(C.byref(g2), fp(mat))
        return g2

    def histogram(self, g, d, cache, nbins):
        d = f32(d)
        bins = (C.c_long * nbins)()
        mid = np.zeros(nbins, np.float32)
        self.lib.vo_histogram(C.byref(g), fp(d), C.byref(cache), nbins, bins, fp(mid))
        return np.array(list(bins), dtype=np.int64), mid

    # ---- resizing ----------------------------------------------------------------------------
    def _resize(self, fn, g, d, *params):
        d = f32(d)
        g2, out = g.copy(), c_float_p()
        fn(C.byref(g2), fp(d), C.byref(out), *params)
        return g2, self._take(out, g2.n)

    def pad(self, g, d, pads):
        return self._resize(self.lib.vo_pad, g, d, *[int(p) for p in pads])

    def crop(self, g, d, box):
        return self._resize(self.lib.vo_crop, g, d, *[float(p) for p in box])

    def downsample(self, g, d):
        return self._resize(self.lib.vo_downsample, g, d)

    def supersample(self, g, d):
        return self._resize(self.lib.vo_supersample, g, d)

    def blur(self, g, d, sigma):
        d = f32(d)
        out = c_float_p()
        self.lib.vo_gaussian_blur(C.byref(g), fp(d), C.byref(out), float(sigma))
        return self._take(out, g.n)

    def blur_kernel(self, sigma):
        w = np.zeros(1024, np.float32)
        R = self.lib.vo_blur_kernel(float(sigma), fp(w), 1024)
        return R, w[:2 * R + 1].copy()

    # ---- sim / fit -----------------------------------------------------------------------------
    def sim_policy(self, resolution, gspacing=0.0):
        rs, gs, q = C.c_float(), C.c_double(), C.c_int()
        self.lib.vo_sim_policy(resolution, gspacing, C.byref(rs), C.byref(gs), C.byref(q))
        return rs.value, gs.value, q.value

    def sim(self, pos, radii, quality, radscale, gspacing, nthreads=1):
        pos, radii = f32(pos), f32(radii)
        g, out = Grid(), c_float_p()
        self.lib.vo_sim(fp(pos), fp(radii), len(radii), quality, radscale, gspacing, C.byref(g), C.byref(out), nthreads)
        return g, self._take(out, g.n)

    def measure_center(self, pos, w):
        pos, w = f32(pos), f32(w)
        com = np.zeros(3, np.float32)
        self.lib.vo_measure_center(fp(pos), fp(w), len(w), fp(com))
        return com

    def pose_transform(self, pos, com, stride, x, y, z):
        p = f32(pos).copy()
        com = f32(com)
        self.lib.vo_pose_transform(fp(p), p.size // 3, fp(com), stride, x, y, z)
        return p

    def pose_general(self, pos, com, rot_deg, shift=(0.0, 0.0, 0.0)):
        """one candidate of an explicit pose list: float degrees per axis, then a translation"""
        p = f32(pos).copy()
        com, rot, sh = f32(com), f32(rot_deg), f32(shift)
        self.lib.vo_pose_general(fp(p), p.size // 3, fp(com), fp(rot), fp(sh))
        return p

    def calc_cc(self, pos, radii, gT, t, resolution, nthreads=1):
        pos, radii, t = f32(pos), f32(radii), f32(t)
        return self.lib.vo_calc_cc(fp(pos), fp(radii), len(radii), C.byref(gT), fp(t), resolution, nthreads)

    def rotate_table(self, origin_pos, radii, com, stride, max_rot, gT, t, resolution, nthreads=1, pose_first=0,
                     pose_stride=1):
        origin_pos, radii, t, com = f32(origin_pos), f32(radii), f32(t), f32(com)
        tab = np.full(max_rot ** 3, np.nan, np.float32)
        self.lib.vo_rotate_table(fp(origin_pos), fp(radii), len(radii), fp(com), stride, max_rot, C.byref(gT), fp(t),
                                 resolution, fp(tab), pose_stride, pose_first, nthreads)
        return tab

    def rotate_replay(self, table, max_rot, returncc):
        table = f32(table)
        cc, bi = C.c_float(returncc), C.c_int(-1)
        visited = self.lib.vo_rotate_replay(fp(table), max_rot, C.byref(cc), C.byref(bi))
        return cc.value, bi.value, visited

    def fit(self, pos, radii, mass, gT, t, resolution, direct=False, nthreads=1):
        p, radii, mass, t = f32(pos).copy(), f32(radii), f32(mass), f32(t)
        res = FitResult()
        self.lib.vo_fit(fp(p), fp(radii), fp(mass), len(radii), C.byref(gT), fp(t), resolution, C.byref(res),
                        int(direct), nthreads)
        return p, res

    # ---- geometry ------------------------------------------------------------------------------
    def mat4_inverse(self, m):
        m = f32(m).copy()
        rc = self.lib.vo_mat4_inverse(fp(m))
        return rc, m

    def grid_inverse(self, g, shift):
        m = np.zeros(16, np.float32)
        self.lib.vo_grid_inverse(C.byref(g), shift, fp(m))
        return m

    def voxel_coord(self, g, i):
        x, y, z = C.c_float(), C.c_float(), C.c_float()
        self.lib.vo_voxel_coord(C.byref(g), i, C.byref(x), C.byref(y), C.byref(z))
        return x.value, y.value, z.value

    def voxel_coord_int(self, g, i):
        x, y, z = C.c_float(), C.c_float(), C.c_float()
        self.lib.vo_voxel_coord_int(C.byref(g), i, C.byref(x), C.byref(y), C.byref(z))
        return x.value, y.value, z.value

    def lookup(self, g, d, mode, xyz):
        """mode: 0 nearest, 1 nearest safe, 2 interp, 3 interp safe (same codes as ref_lookup)"""
        d, xyz = f32(d), f32(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), np.float32)
        for i, (x, y, z) in enumerate(xyz):
            if mode < 2:
                out[i] = self.lib.vo_value_from_coord(C.byref(g), fp(d), x, y, z, mode & 1)
            else:
                out[i] = self.lib.vo_value_interp_from_coord(C.byref(g), fp(d), x, y, z, mode & 1)
        return out

    def intersection(self, gA, gB):
        o = Grid()
        self.lib.vo_init_from_intersection(C.byref(gA), C.byref(gB), C.byref(o))
        return o

    def union(self, gA, gB):
        o = Grid()
        self.lib.vo_init_from_union(C.byref(gA), C.byref(gB), C.byref(o))
        return o


class Ref(Oracle):
    """the reference's own object code (oracle/_ref); superset of Oracle (vo_* symbols included)."""
    LIB = os.path.join(HERE, "_ref", "libvoltool_ref.so")
    OPS = {"clamp": 0, "scale_by": 1, "scalar_add": 2, "rescale_range": 3, "sigma_scale": 4, "binmask": 5,
           "mdff_potential": 6, "downsample": 7, "supersample": 8, "blur": 9, "pad": 10, "crop": 11}

    def __init__(self, path=None):
        """path: another library with the same flat API over the reference's classes -- the compiled adapter
        shim/_build/libvoltool_shim.so runs the SAME VolumetricData class with GPU bodies"""
        if path is None and not os.path.exists(self.LIB):
            build(ref=True)
        super().__init__(path or self.LIB)
        L = self.lib
        L.ref_vol_new.argtypes = [C.POINTER(Grid), c_float_p]
        L.ref_vol_new.restype = C.c_void_p
        L.ref_vol_free.argtypes = [C.c_void_p]
        L.ref_vol_grid.argtypes = [C.c_void_p, C.POINTER(Grid)]
        L.ref_vol_data.argtypes = [C.c_void_p]
        L.ref_vol_data.restype = c_float_p
        L.ref_vol_cache.argtypes = [C.c_void_p, C.POINTER(Cache)]
        L.ref_vol_op.argtypes = [C.c_void_p, C.c_int, c_double_p]
        L.ref_vol_datarange.argtypes = [C.c_void_p, c_float_p, c_float_p]
        L.ref_vol_mean.argtypes = [C.c_void_p]
        L.ref_vol_mean.restype = C.c_float
        L.ref_vol_sigma.argtypes = [C.c_void_p]
        L.ref_vol_sigma.restype = C.c_float
        L.ref_vol_com.argtypes = [C.c_void_p, c_float_p]
        L.ref_vol_moveto.argtypes = [C.c_void_p, c_float_p, c_float_p]
        L.ref_vol_move.argtypes = [C.c_void_p, c_float_p]
        L.ref_vol_histogram.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_long), c_float_p]
        L.ref_voxel_coord.argtypes = [C.c_void_p, C.c_long, c_float_p]
        L.ref_voxel_coord_int.argtypes = [C.c_void_p, C.c_int, c_float_p]
        L.ref_voxel_coord_from_cartesian.argtypes = [C.c_void_p, c_float_p, c_float_p, C.c_int]
        L.ref_voxel_index_from_coord.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.ref_voxel_index_from_coord.restype = C.c_long
        L.ref_lookup.argtypes = [C.c_void_p, C.c_int, C.c_long, c_float_p, c_float_p]
        L.ref_init_from_intersection.argtypes = [C.POINTER(Grid)] * 3
        L.ref_init_from_union.argtypes = [C.POINTER(Grid)] * 3
        L.ref_cc.argtypes = [C.POINTER(Grid), c_float_p, C.POINTER(Grid), c_float_p, C.c_double, c_double_p]
        L.ref_cc_vols.argtypes = [C.c_void_p, C.c_void_p, C.c_double, c_double_p]
        L.ref_binary.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(Grid), c_float_p, C.POINTER(Grid), c_float_p,
                                 C.POINTER(Grid), C.POINTER(c_float_p)]
        L.ref_install_cc_backend.argtypes = [C.c_void_p]
        L.ref_set_numprocs.argtypes = [C.c_int]

    class Vol:
        """a live reference VolumetricData object"""

        def __init__(self, ref, g, d):
            self.ref = ref
            d = f32(d)
            self.h = ref.lib.ref_vol_new(C.byref(g), fp(d))

        def close(self):
            if self.h:
                self.ref.lib.ref_vol_free(self.h)
                self.h = None

        def __del__(self):
            self.close()

        @property
        def grid(self):
            g = Grid()
            self.ref.lib.ref_vol_grid(self.h, C.byref(g))
            return g

        @property
        def data(self):
            g = self.grid
            return np.ctypeslib.as_array(self.ref.lib.ref_vol_data(self.h), shape=(max(g.n, 1),))[:g.n].copy()

        @property
        def cache(self):
            c = Cache()
            self.ref.lib.ref_vol_cache(self.h, C.byref(c))
            return c

        def op(self, name, *params):
            p = (C.c_double * 6)(*([float(x) for x in params] + [0.0] * (6 - len(params))))
            self.ref.lib.ref_vol_op(self.h, Ref.OPS[name], p)
            return self

        def datarange(self):
            mn, mx = C.c_float(), C.c_float()
            self.ref.lib.ref_vol_datarange(self.h, C.byref(mn), C.byref(mx))
            return mn.value, mx.value

        def mean(self):
            return self.ref.lib.ref_vol_mean(self.h)

        def sigma(self):
            return self.ref.lib.ref_vol_sigma(self.h)

        def com(self):
            com = np.zeros(3, np.float32)
            self.ref.lib.ref_vol_com(self.h, fp(com))
            return com

        def moveto(self, com, pos):
            com, pos = f32(com), f32(pos)
            self.ref.lib.ref_vol_moveto(self.h, fp(com), fp(pos))
            return self

        def move(self, mat):
            mat = f32(mat)
            self.ref.lib.ref_vol_move(self.h, fp(mat))
            return self

        def histogram(self, nbins):
            bins = (C.c_long * nbins)()
            mid = np.zeros(nbins, np.float32)
            self.ref.lib.ref_vol_histogram(self.h, nbins, bins, fp(mid))
            return np.array(list(bins), dtype=np.int64), mid

        def voxel_coord(self, i):
            xyz = np.zeros(3, np.float32)
            self.ref.lib.ref_voxel_coord(self.h, i, fp(xyz))
            return tuple(xyz)

        def voxel_coord_int(self, i):
            xyz = np.zeros(3, np.float32)
            self.ref.lib.ref_voxel_coord_int(self.h, i, fp(xyz))
            return tuple(xyz)

        def lookup(self, mode, xyz):
            xyz = f32(xyz).reshape(-1, 3)
            out = np.zeros(len(xyz), np.float32)
            self.ref.lib.ref_lookup(self.h, mode, len(xyz), fp(xyz), fp(out))
            return out

    def vol(self, g, d):
        return Ref.Vol(self, g, d)

    def set_numprocs(self, n):
        self.lib.ref_set_numprocs(n)

    def ref_cc(self, gA, a, gB, b, thr=-3.4028234663852886e38):
        a, b = f32(a), f32(b)
        cc = C.c_double()
        self.lib.ref_cc(C.byref(gA), fp(a), C.byref(gB), fp(b), thr, C.byref(cc))
        return cc.value

    def ref_cc_vols(self, va, vb, thr):
        cc = C.c_double()
        self.lib.ref_cc_vols(va.h, vb.h, thr, C.byref(cc))
        return cc.value

    def ref_binary(self, op, gA, a, gB, b, interp=True, use_union=False):
        a, b = f32(a), f32(b)
        og, out = Grid(), c_float_p()
        self.lib.ref_binary(op, int(interp), int(use_union), C.byref(gA), fp(a), C.byref(gB), fp(b), C.byref(og),
                            C.byref(out))
        return og, self._take(out, og.n)

    def ref_intersection(self, gA, gB):
        o = Grid()
        self.lib.ref_init_from_intersection(C.byref(gA), C.byref(gB), C.byref(o))
        return o

    def ref_union(self, gA, gB):
        o = Grid()
        self.lib.ref_init_from_union(C.byref(gA), C.byref(gB), C.byref(o))
        return o

    def install_cc_backend(self, target_vol):
        """route the restated fit driver's CC through the reference's cc_threaded"""
        self.lib.ref_install_cc_backend(target_vol.h if target_vol is not None else None)
