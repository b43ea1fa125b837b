"""ctypes face of libvoltool_gpu.so (include/voltool_gpu.h).

Names follow the reference: ``VolMap`` carries the VolumetricData methods (VolumetricData.h),
module functions carry the free functions (cc_threaded MDFF.C:102, add/subtract/multiply/average
Voltool.C:249-377, fit TclVoltool.C:679-928).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FLT_MAX = 3.4028234663852886e38
c_float_p = C.POINTER(C.c_float)


class VoltoolGpuError(RuntimeError):
    pass


class Grid(C.Structure):
    """vt_grid: mirrors VolumetricData.h:29-33"""
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

    @classmethod
    def from_fields(cls, other):
        """copy from any ctypes struct with the same fields (e.g. the checker's Grid)"""
        g = cls()
        for k in range(3):
            g.origin[k], g.xaxis[k], g.yaxis[k], g.zaxis[k] = other.origin[k], other.xaxis[k], other.yaxis[k], other.zaxis[k]
        g.xsize, g.ysize, g.zsize = other.xsize, other.ysize, other.zsize
        return g

    @property
    def n(self):
        return self.xsize * self.ysize * self.zsize

    @property
    def shape(self):
        return (self.zsize, self.ysize, self.xsize)

    def astuple(self):
        return (tuple(self.origin), tuple(self.xaxis), tuple(self.yaxis), tuple(self.zaxis),
                (self.xsize, self.ysize, self.zsize))


class Cache(C.Structure):
    """vt_cache: VolumetricData.h:202-210"""
    _fields_ = [("minmax_valid", C.c_int), ("mean_valid", C.c_int), ("sigma_valid", C.c_int),
                ("cmin", C.c_float), ("cmax", C.c_float), ("cmean", C.c_float), ("csigma", C.c_float)]

    def astuple(self):
        return (self.minmax_valid, self.mean_valid, self.sigma_valid, self.cmin, self.cmax, self.cmean, self.csigma)


class Pose(C.Structure):
    _fields_ = [("rot_deg", C.c_float * 3), ("shift", C.c_float * 3)]


class FitResult(C.Structure):
    _fields_ = [("best_cc", C.c_float), ("com", C.c_float * 3), ("dencom", C.c_float * 3), ("n_evaluated", C.c_int),
                ("n_computed", C.c_int), ("stage_best", C.c_int * 5), ("exhaustive_cc", C.c_float)]


MERGE_FN = C.CFUNCTYPE(C.c_int, c_float_p, C.c_int, C.c_int, C.c_int, C.c_void_p)
EVAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p)

_lib = None
_inited = False


def lib_path():
    return os.environ.get("VOLTOOL_GPU_LIB", os.path.join(HERE, "libvoltool_gpu.so"))


def lib():
    """dlopen the product library; raises (no fallback) when it has not been built"""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise VoltoolGpuError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C voltool_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(path)
    G, M = C.POINTER(Grid), C.c_void_p
    L.vt_last_error.restype = C.c_char_p
    L.vt_init.argtypes = [C.c_int]
    L.vt_launch_count.restype = C.c_long
    L.vt_stream.restype = C.c_void_p
    L.vt_profile_enable.argtypes = [C.c_int]
    L.vt_profile_get.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_long)]
    L.vt_device_props.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.c_char_p, C.c_int]
    L.vt_map_create.argtypes = [G, C.c_void_p, C.POINTER(M)]
    L.vt_map_create_from_device.argtypes = [G, C.c_void_p, C.POINTER(M)]
    L.vt_map_clone.argtypes = [M, C.POINTER(M)]
    L.vt_map_destroy.argtypes = [M]
    L.vt_map_grid.argtypes = [M, G]
    L.vt_map_cache.argtypes = [M, C.POINTER(Cache)]
    L.vt_map_gridsize.argtypes = [M]
    L.vt_map_gridsize.restype = C.c_long
    L.vt_map_download.argtypes = [M, C.c_void_p]
    L.vt_map_upload.argtypes = [M, C.c_void_p]
    L.vt_map_device_ptr.argtypes = [M]
    L.vt_map_device_ptr.restype = C.c_void_p
    L.vt_cc.argtypes = [M, M, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_double)]
    L.vt_cc_host.argtypes = [G, C.c_void_p, G, C.c_void_p, C.c_double, C.POINTER(C.c_double)]
    for name in ("vt_clamp", "vt_rescale_voxel_value_range"):
        getattr(L, name).argtypes = [M, C.c_float, C.c_float]
    for name in ("vt_scale_by", "vt_scalar_add", "vt_binmask"):
        getattr(L, name).argtypes = [M, C.c_float]
    L.vt_sigma_scale.argtypes = [M]
    L.vt_mdff_potential.argtypes = [M, C.c_double]
    L.vt_datarange.argtypes = [M, c_float_p, c_float_p]
    L.vt_mean.argtypes = [M, c_float_p]
    L.vt_sigma.argtypes = [M, c_float_p]
    L.vt_vol_com.argtypes = [M, c_float_p]
    L.vt_histogram.argtypes = [M, C.c_int, C.POINTER(C.c_long), c_float_p]
    L.vt_vol_moveto.argtypes = [M, c_float_p, c_float_p]
    L.vt_vol_move.argtypes = [M, c_float_p]
    L.vt_pad.argtypes = [M] + [C.c_int] * 6
    L.vt_crop.argtypes = [M] + [C.c_double] * 6
    L.vt_downsample.argtypes = [M]
    L.vt_supersample.argtypes = [M]
    L.vt_gaussian_blur.argtypes = [M, C.c_double]
    L.vt_binary.argtypes = [C.c_int, C.c_int, C.c_int, M, M, C.POINTER(M)]
    L.vt_binary_host.argtypes = [C.c_int, C.c_int, C.c_int, G, C.c_void_p, G, C.c_void_p, G, C.POINTER(c_float_p)]
    L.vt_free_host.argtypes = [C.c_void_p]
    L.vt_dx_write.argtypes = [C.c_char_p, G, C.c_void_p, C.c_char_p]
    L.vt_dx_read.argtypes = [C.c_char_p, G, C.POINTER(c_float_p)]
    L.vt_map_write_dx.argtypes = [M, C.c_char_p, C.c_char_p]
    L.vt_map_read_dx.argtypes = [C.c_char_p, C.POINTER(M)]
    L.vt_mrc_write.argtypes = [C.c_char_p, G, C.c_void_p]
    L.vt_mrc_read.argtypes = [C.c_char_p, G, C.POINTER(c_float_p)]
    L.vt_map_write_mrc.argtypes = [M, C.c_char_p]
    L.vt_map_read_mrc.argtypes = [C.c_char_p, C.POINTER(M)]
    L.vt_init_from_intersection.argtypes = [G, G, G]
    L.vt_init_from_union.argtypes = [G, G, G]
    L.vt_sim_policy.argtypes = [C.c_double, C.c_double, c_float_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.vt_sim.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_float, C.c_float, C.POINTER(M)]
    L.vt_calc_cc.argtypes = [C.c_void_p, C.c_void_p, C.c_long, M, C.c_float, C.POINTER(C.c_double)]
    L.vt_sim_cc.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_double, C.c_double, M, C.c_double,
                            C.POINTER(C.c_double), C.POINTER(M)]
    L.vt_fit_create.argtypes = [C.c_void_p, C.c_void_p, C.c_long, M, C.c_float, C.POINTER(C.c_void_p)]
    L.vt_fit_destroy.argtypes = [C.c_void_p]
    L.vt_fit_set_origin.argtypes = [C.c_void_p, C.c_void_p]
    L.vt_fit_eval_poses.argtypes = [C.c_void_p, c_float_p, C.c_void_p, C.c_long, C.c_void_p]
    L.vt_fit_eval_poses_device.argtypes = [C.c_void_p, c_float_p, C.c_void_p, C.c_long, C.c_void_p]
    L.vt_fit_rotate_table.argtypes = [C.c_void_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.vt_rotate_replay.argtypes = [C.c_void_p, C.c_int, c_float_p, C.POINTER(C.c_int)]
    L.vt_pose_apply.argtypes = [C.c_void_p, C.c_long, c_float_p, C.POINTER(Pose)]
    L.vt_measure_center.argtypes = [C.c_void_p, C.c_void_p, C.c_long, c_float_p]
    L.vt_fit_set_shard.argtypes = [C.c_int, C.c_int, MERGE_FN, C.c_void_p]
    L.vt_fit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, M, C.c_float, C.POINTER(FitResult)]
    L.vt_fit_schedule.argtypes = [C.c_void_p, C.c_void_p, C.c_long, c_float_p, EVAL_FN, C.c_void_p, C.POINTER(FitResult)]
    IP, LP = C.POINTER(C.c_int), C.POINTER(C.c_long)
    L.vt_compact_selection.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, LP]
    L.vt_sim_sel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_float, C.c_float, C.POINTER(M)]
    L.vt_sim_cc_sel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_double, C.c_double, M, C.c_double,
                                C.POINTER(C.c_double), C.POINTER(M)]
    L.vt_fit_sel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, M, C.c_float, C.POINTER(FitResult)]
    L.vt_fit_info.argtypes = [C.c_void_p, IP, IP, IP, IP]
    L.vt_nccl_unique_id.argtypes = [C.c_void_p, C.c_int]
    L.vt_nccl_init.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.vt_fit_best_pose_device.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_void_p]
    L.vt_nccl_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    L.vt_cc_maps.argtypes = [M, M, C.c_double, C.POINTER(C.c_double), C.POINTER(M), C.POINTER(M)]
    L.vt_sim_cc_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_double, C.c_double, M, C.c_double,
                                 C.POINTER(C.c_double), C.POINTER(M), C.POINTER(M), C.POINTER(M)]
    L.vt_density_rotate.argtypes = [M, M, C.c_int, C.c_int, c_float_p, c_float_p, IP, C.c_void_p]
    L.vt_density_reset.argtypes = [M, C.c_int, IP, c_float_p, c_float_p]
    L.vt_fit_search_grid.argtypes = [C.c_void_p, c_float_p, C.c_void_p, C.c_long, C.c_int, C.c_float, c_float_p, LP,
                                     C.POINTER(Pose), C.c_void_p]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise VoltoolGpuError(f"{what} failed ({rc}): {lib().vt_last_error().decode(errors='replace')}")


def init(device=-1):
    """vt_init: raises when no CUDA device is usable (the product never falls back to the CPU)"""
    global _inited
    _check(lib().vt_init(int(device)), "vt_init")
    _inited = True


def _need():
    if not _inited:
        init()


def shutdown():
    global _inited
    if _lib is not None:
        _lib.vt_shutdown()
    _inited = False


def sync():
    _need()
    _check(lib().vt_sync(), "vt_sync")


def launch_count():
    return lib().vt_launch_count()


def device_props():
    _need()
    sm, khz, mem = C.c_int(), C.c_int(), C.c_size_t()
    name = C.create_string_buffer(128)
    _check(lib().vt_device_props(C.byref(sm), C.byref(khz), C.byref(mem), name, 128), "vt_device_props")
    return {"sm_count": sm.value, "clock_khz": khz.value, "mem_bytes": mem.value, "name": name.value.decode()}


def profile_enable(on=True):
    lib().vt_profile_enable(int(on))


def profile_reset():
    lib().vt_profile_reset()


def profile_get(kernel_class):
    """(total ms, launches) of one kernel class since the last reset (CUDA events on the library stream)"""
    ms, n = C.c_double(), C.c_long()
    _check(lib().vt_profile_get(kernel_class.encode(), C.byref(ms), C.byref(n)), "vt_profile_get")
    return ms.value, n.value


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _fp3(v):
    a = (C.c_float * 3)(*[float(x) for x in v])
    return C.cast(a, c_float_p), a


class VolMap:
    """resident twin of VolumetricData (VolumetricData.h:27-36); methods keep the reference's names"""

    def __init__(self, grid=None, data=None, _handle=None):
        _need()
        if _handle is not None:
            self.h = _handle
            return
        h = C.c_void_p()
        g = grid if isinstance(grid, Grid) else Grid.from_fields(grid)
        if data is not None:
            data = f32(data).reshape(-1)
            if data.size != g.n:
                raise ValueError("data size does not match grid")
            _check(lib().vt_map_create(C.byref(g), _vp(data), C.byref(h)), "vt_map_create")
        else:
            _check(lib().vt_map_create(C.byref(g), None, C.byref(h)), "vt_map_create")
        self.h = h

    @classmethod
    def from_device(cls, grid, dev_ptr):
        _need()
        h = C.c_void_p()
        _check(lib().vt_map_create_from_device(C.byref(grid), C.c_void_p(dev_ptr), C.byref(h)), "vt_map_create_from_device")
        return cls(_handle=h)

    def close(self):
        if getattr(self, "h", None) and _lib is not None and _inited:
            _lib.vt_map_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clone(self):
        h = C.c_void_p()
        _check(lib().vt_map_clone(self.h, C.byref(h)), "vt_map_clone")
        return VolMap(_handle=h)

    # ---- accessors -----------------------------------------------------------------------------
    @property
    def grid(self):
        g = Grid()
        lib().vt_map_grid(self.h, C.byref(g))
        return g

    @property
    def cache(self):
        c = Cache()
        lib().vt_map_cache(self.h, C.byref(c))
        return c

    def gridsize(self):
        return lib().vt_map_gridsize(self.h)

    @property
    def device_ptr(self):
        return lib().vt_map_device_ptr(self.h)

    @property
    def data(self):
        out = np.empty(max(self.gridsize(), 0), np.float32)
        if out.size:
            _check(lib().vt_map_download(self.h, _vp(out)), "vt_map_download")
        return out

    def upload(self, data):
        data = f32(data).reshape(-1)
        _check(lib().vt_map_upload(self.h, _vp(data)), "vt_map_upload")

    # ---- VolumetricData methods ------------------------------------------------------------------
    def write_dx(self, path, name="density"):
        _need()
        _check(lib().vt_map_write_dx(self.h, os.fsencode(path), name.encode()), "vt_map_write_dx")

    @classmethod
    def read_dx(cls, path):
        _need()
        h = C.c_void_p()
        _check(lib().vt_map_read_dx(os.fsencode(path), C.byref(h)), "vt_map_read_dx")
        return cls(_handle=h)

    def write_mrc(self, path):
        _need()
        _check(lib().vt_map_write_mrc(self.h, os.fsencode(path)), "vt_map_write_mrc")

    @classmethod
    def read_mrc(cls, path):
        _need()
        h = C.c_void_p()
        _check(lib().vt_map_read_mrc(os.fsencode(path), C.byref(h)), "vt_map_read_mrc")
        return cls(_handle=h)

    @classmethod
    def read_file(cls, path):
        """by extension: .mrc / .map / .ccp4 -> MRC, anything else -> OpenDX"""
        return cls.read_mrc(path) if is_mrc_path(path) else cls.read_dx(path)

    def write_file(self, path, name="density"):
        if is_mrc_path(path):
            self.write_mrc(path)
        else:
            self.write_dx(path, name)

    def clamp(self, lo, hi):
        _check(lib().vt_clamp(self.h, lo, hi), "vt_clamp"); return self

    def scale_by(self, ff):
        _check(lib().vt_scale_by(self.h, ff), "vt_scale_by"); return self

    def scalar_add(self, ff):
        _check(lib().vt_scalar_add(self.h, ff), "vt_scalar_add"); return self

    def rescale_voxel_value_range(self, lo, hi):
        _check(lib().vt_rescale_voxel_value_range(self.h, lo, hi), "vt_rescale_voxel_value_range"); return self

    def sigma_scale(self):
        _check(lib().vt_sigma_scale(self.h), "vt_sigma_scale"); return self

    def binmask(self, threshold=0.0):
        _check(lib().vt_binmask(self.h, threshold), "vt_binmask"); return self

    def mdff_potential(self, threshold):
        _check(lib().vt_mdff_potential(self.h, threshold), "vt_mdff_potential"); return self

    def datarange(self):
        a, b = C.c_float(), C.c_float()
        _check(lib().vt_datarange(self.h, C.byref(a), C.byref(b)), "vt_datarange")
        return a.value, b.value

    def mean(self):
        a = C.c_float()
        _check(lib().vt_mean(self.h, C.byref(a)), "vt_mean")
        return a.value

    def sigma(self):
        a = C.c_float()
        _check(lib().vt_sigma(self.h, C.byref(a)), "vt_sigma")
        return a.value

    def vol_com(self):
        com = np.zeros(3, np.float32)
        _check(lib().vt_vol_com(self.h, com.ctypes.data_as(c_float_p)), "vt_vol_com")
        return com

    def vol_moveto(self, com, pos):
        a, ka = _fp3(com)
        b, kb = _fp3(pos)
        _check(lib().vt_vol_moveto(self.h, a, b), "vt_vol_moveto"); return self

    def vol_move(self, mat):
        m = f32(mat).reshape(-1)
        _check(lib().vt_vol_move(self.h, m.ctypes.data_as(c_float_p)), "vt_vol_move"); return self

    def histogram(self, nbins=10):
        bins = (C.c_long * nbins)()
        mid = np.zeros(nbins, np.float32)
        _check(lib().vt_histogram(self.h, nbins, bins, mid.ctypes.data_as(c_float_p)), "vt_histogram")
        return np.array(list(bins), dtype=np.int64), mid

    def pad(self, pxm, pxp, pym, pyp, pzm, pzp):
        _check(lib().vt_pad(self.h, pxm, pxp, pym, pyp, pzm, pzp), "vt_pad"); return self

    def crop(self, minx, miny, minz, maxx, maxy, maxz):
        _check(lib().vt_crop(self.h, minx, miny, minz, maxx, maxy, maxz), "vt_crop"); return self

    def downsample(self):
        _check(lib().vt_downsample(self.h), "vt_downsample"); return self

    def supersample(self):
        _check(lib().vt_supersample(self.h), "vt_supersample"); return self

    def gaussian_blur(self, sigma):
        _check(lib().vt_gaussian_blur(self.h, sigma), "vt_gaussian_blur"); return self


# ---- free functions ------------------------------------------------------------------------------
def write_dx(path, grid, data, name="density"):
    """host arrays -> OpenDX file (no GPU needed); VMD volmap layout, TclMDFF.C:159"""
    d = f32(data)
    if d.size != grid.n:
        raise ValueError("data size does not match the grid")
    _check(lib().vt_dx_write(os.fsencode(path), C.byref(grid), _vp(d), name.encode()), "vt_dx_write")


def read_dx(path):
    """OpenDX file -> (Grid, float32 array in map order: x fastest); no GPU needed"""
    g = Grid()
    p = c_float_p()
    _check(lib().vt_dx_read(os.fsencode(path), C.byref(g), C.byref(p)), "vt_dx_read")
    try:
        out = np.ctypeslib.as_array(p, shape=(max(g.n, 1),))[:g.n].copy()
    finally:
        lib().vt_free_host(p)
    return g, out


def is_mrc_path(path):
    return os.fspath(path).lower().endswith((".mrc", ".map", ".ccp4"))


def write_mrc(path, grid, data):
    """host arrays -> MRC2014 file, mode 2 (no GPU needed)"""
    d = f32(data)
    if d.size != grid.n:
        raise ValueError("data size does not match the grid")
    _check(lib().vt_mrc_write(os.fsencode(path), C.byref(grid), _vp(d)), "vt_mrc_write")


def read_mrc(path):
    """MRC / CCP4 file -> (Grid, float32 array in map order: x fastest); no GPU needed"""
    g = Grid()
    p = c_float_p()
    _check(lib().vt_mrc_read(os.fsencode(path), C.byref(g), C.byref(p)), "vt_mrc_read")
    try:
        out = np.ctypeslib.as_array(p, shape=(max(g.n, 1),))[:g.n].copy()
    finally:
        lib().vt_free_host(p)
    return g, out


def cc_threaded(qs_vol, target_vol, threshold=-FLT_MAX, full=False):
    """cc_threaded(qsVol, targetVol, &cc, threshold)  MDFF.C:102"""
    cc, n, sums = C.c_double(), C.c_long(), (C.c_double * 5)()
    _check(lib().vt_cc(qs_vol.h, target_vol.h, threshold, C.byref(cc), C.byref(n), sums), "vt_cc")
    return (cc.value, n.value, list(sums)) if full else cc.value


def cc_host(grid_a, a, grid_b, b, threshold=-FLT_MAX):
    """vt_cc_host: the body INTEGRATION.md gives cc_threaded (host arrays in, cc out)"""
    _need()
    a, b = f32(a).reshape(-1), f32(b).reshape(-1)
    cc = C.c_double()
    _check(lib().vt_cc_host(C.byref(grid_a), _vp(a), C.byref(grid_b), _vp(b), threshold, C.byref(cc)), "vt_cc_host")
    return cc.value


def binary_host(op, grid_a, a, grid_b, b, interp=True, use_union=False):
    """vt_binary_host: host arrays in, (Grid, float32 array) out"""
    _need()
    a, b = f32(a).reshape(-1), f32(b).reshape(-1)
    g, p = Grid(), c_float_p()
    _check(lib().vt_binary_host(op, int(interp), int(use_union), C.byref(grid_a), _vp(a), C.byref(grid_b), _vp(b),
                                C.byref(g), C.byref(p)), "vt_binary_host")
    try:
        out = np.ctypeslib.as_array(p, shape=(max(g.n, 1),))[:g.n].copy()
    finally:
        lib().vt_free_host(p)
    return g, out


def binary(op, map_a, map_b, interp=True, use_union=False):
    h = C.c_void_p()
    _check(lib().vt_binary(op, int(interp), int(use_union), map_a.h, map_b.h, C.byref(h)), "vt_binary")
    return VolMap(_handle=h)


def add(a, b, interp=True, use_union=False):
    return binary(0, a, b, interp, use_union)


def subtract(a, b, interp=True, use_union=False):
    return binary(1, a, b, interp, use_union)


def multiply(a, b, interp=True, use_union=False):
    return binary(2, a, b, interp, use_union)


def average(a, b, interp=True, use_union=False):
    return binary(3, a, b, interp, use_union)


def init_from_intersection(a, b):
    o = Grid()
    lib().vt_init_from_intersection(C.byref(a), C.byref(b), C.byref(o))
    return o


def init_from_union(a, b):
    o = Grid()
    lib().vt_init_from_union(C.byref(a), C.byref(b), C.byref(o))
    return o


def sim_policy(resolution, gspacing=0.0):
    rs, gs, q = C.c_float(), C.c_double(), C.c_int()
    lib().vt_sim_policy(resolution, gspacing, C.byref(rs), C.byref(gs), C.byref(q))
    return rs.value, gs.value, q.value


def sim(pos, radii, quality, radscale, gspacing):
    """QuickSurf::calc_density_map call sites TclMDFF.C:154-157"""
    _need()
    pos, radii = f32(pos).reshape(-1), f32(radii).reshape(-1)
    h = C.c_void_p()
    _check(lib().vt_sim(_vp(pos), _vp(radii), radii.size, quality, radscale, gspacing, C.byref(h)), "vt_sim")
    return VolMap(_handle=h)


def calc_cc(pos, radii, target, resolution):
    """calc_cc, TclVoltool.C:73-129"""
    _need()
    pos, radii = f32(pos).reshape(-1), f32(radii).reshape(-1)
    cc = C.c_double()
    _check(lib().vt_calc_cc(_vp(pos), _vp(radii), radii.size, target.h, resolution, C.byref(cc)), "vt_calc_cc")
    return cc.value


def sim_cc(pos, radii, resolution, target, threshold=-FLT_MAX, gspacing=0.0, want_synth=False):
    """mdff_cc CPU branch, TclMDFF.C:486-524"""
    _need()
    pos, radii = f32(pos).reshape(-1), f32(radii).reshape(-1)
    cc, h = C.c_double(), C.c_void_p()
    _check(lib().vt_sim_cc(_vp(pos), _vp(radii), radii.size, resolution, gspacing, target.h, threshold, C.byref(cc),
                           C.byref(h) if want_synth else None), "vt_sim_cc")
    return (cc.value, VolMap(_handle=h)) if want_synth else cc.value


def cc_maps(map_a, map_b, threshold=-FLT_MAX, want_diff=True, want_spatial=True):
    """cc + the -calcdiffmap / -calcspatialccmap outputs (TclMDFF.C:180-184, 555-594): (cc, diff, spatial)"""
    _need()
    cc, hd, hs = C.c_double(), C.c_void_p(), C.c_void_p()
    _check(lib().vt_cc_maps(map_a.h, map_b.h, threshold, C.byref(cc), C.byref(hd) if want_diff else None,
                            C.byref(hs) if want_spatial else None), "vt_cc_maps")
    return cc.value, (VolMap(_handle=hd) if want_diff else None), (VolMap(_handle=hs) if want_spatial else None)


def sim_cc_maps(pos, radii, resolution, target, threshold=-FLT_MAX, gspacing=0.0):
    """mdff_cc with -calcsynthmap -calcdiffmap -calcspatialccmap: (cc, synth, diff, spatial)"""
    _need()
    pos, radii = f32(pos).reshape(-1), f32(radii).reshape(-1)
    cc, h0, h1, h2 = C.c_double(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    _check(lib().vt_sim_cc_maps(_vp(pos), _vp(radii), radii.size, resolution, gspacing, target.h, threshold,
                                C.byref(cc), C.byref(h0), C.byref(h1), C.byref(h2)), "vt_sim_cc_maps")
    return cc.value, VolMap(_handle=h0), VolMap(_handle=h1), VolMap(_handle=h2)


def density_rotate(target, synth, stride, max_rot, com, returncc=-1.0):
    """density_rotate, TclVoltool.C:202-279: (returncc, bestrot or None, table of all max_rot^3 scores)"""
    _need()
    cp, keep = _fp3(com)
    rc_ = C.c_float(returncc)
    best = (C.c_int * 3)(-1, -1, -1)
    tab = np.full(max_rot ** 3, np.nan, np.float32)
    _check(lib().vt_density_rotate(target.h, synth.h, stride, max_rot, cp, C.byref(rc_), best, _vp(tab)),
           "vt_density_rotate")
    return rc_.value, (None if best[0] < 0 else tuple(best)), tab


def density_reset(synth, stride, bestrot, synthcom, com):
    """the map half of reset_density, TclVoltool.C:283-300"""
    _need()
    a, ka = _fp3(synthcom)
    b, kb = _fp3(com)
    best = (C.c_int * 3)(*[int(v) for v in bestrot])
    _check(lib().vt_density_reset(synth.h, stride, best, a, b), "vt_density_reset")
    return synth


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32).reshape(-1)


def compact_selection(pos, radii, mass, on):
    """the AtomSel walk of the reference (firstsel..lastsel, on[i]): returns compacted pos, radii, mass, indices"""
    pos, radii, mass, on = f32(pos).reshape(-1), f32(radii).reshape(-1), f32(mass).reshape(-1), _i32(on)
    n = on.size
    po, ro, mo = np.empty(3 * n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
    io, k = np.empty(n, np.int64), C.c_long()
    _check(lib().vt_compact_selection(_vp(pos), _vp(radii), _vp(mass), _vp(on), n, _vp(po), _vp(ro), _vp(mo), _vp(io),
                                      C.byref(k)), "vt_compact_selection")
    k = k.value
    return po[:3 * k].reshape(-1, 3).copy(), ro[:k].copy(), mo[:k].copy(), io[:k].copy()


def sim_sel(pos, radii, on, quality, radscale, gspacing):
    _need()
    pos, radii, on = f32(pos).reshape(-1), f32(radii).reshape(-1), _i32(on)
    h = C.c_void_p()
    _check(lib().vt_sim_sel(_vp(pos), _vp(radii), _vp(on), on.size, quality, radscale, gspacing, C.byref(h)), "vt_sim_sel")
    return VolMap(_handle=h)


def sim_cc_sel(pos, radii, on, resolution, target, threshold=-FLT_MAX, gspacing=0.0):
    _need()
    pos, radii, on = f32(pos).reshape(-1), f32(radii).reshape(-1), _i32(on)
    cc = C.c_double()
    _check(lib().vt_sim_cc_sel(_vp(pos), _vp(radii), _vp(on), on.size, resolution, gspacing, target.h, threshold,
                               C.byref(cc), None), "vt_sim_cc_sel")
    return cc.value


def fit_sel(pos, radii, mass, on, target, resolution):
    """fit() on the selected atoms of the molecule's full arrays; unselected coordinates come back untouched"""
    _need()
    p, radii, mass, on = f32(pos).reshape(-1).copy(), f32(radii).reshape(-1), f32(mass).reshape(-1), _i32(on)
    res = FitResult()
    _check(lib().vt_fit_sel(_vp(p), _vp(radii), _vp(mass), _vp(on), on.size, target.h, resolution, C.byref(res)), "vt_fit_sel")
    return p.reshape(-1, 3), res


def nccl_unique_id():
    """128-byte NCCL unique id (rank 0 makes it; the caller carries it to the other ranks)"""
    buf = C.create_string_buffer(128)
    _check(lib().vt_nccl_unique_id(buf, 128), "vt_nccl_unique_id")
    return buf.raw


def nccl_init(rank, world, uid):
    """bring up the library's own communicator: vt_fit then shards its pose search and merges over NCCL"""
    _need()
    buf = C.create_string_buffer(bytes(uid), 128)
    _check(lib().vt_nccl_init(rank, world, buf, 128), "vt_nccl_init")


def nccl_shutdown():
    if _lib is not None:
        _lib.vt_nccl_shutdown()


def best_pose_device(d_cc_ptr, n, index_base, d_best2_ptr):
    _check(lib().vt_fit_best_pose_device(C.c_void_p(d_cc_ptr), n, index_base, C.c_void_p(d_best2_ptr)), "vt_fit_best_pose_device")


def nccl_allgather(d_send_ptr, d_recv_ptr, count):
    _check(lib().vt_nccl_allgather(C.c_void_p(d_send_ptr), C.c_void_p(d_recv_ptr), count), "vt_nccl_allgather")


def rotate_replay(table, max_rot, returncc):
    table = f32(table)
    cc, bi = C.c_float(returncc), C.c_int(-1)
    visited = lib().vt_rotate_replay(_vp(table), max_rot, C.byref(cc), C.byref(bi))
    return cc.value, bi.value, visited


def pose_apply(pos, com, rot_deg, shift=(0, 0, 0)):
    p = f32(pos).reshape(-1).copy()
    pose = Pose((C.c_float * 3)(*rot_deg), (C.c_float * 3)(*shift))
    cp, keep = _fp3(com)
    lib().vt_pose_apply(_vp(p), p.size // 3, cp, C.byref(pose))
    return p.reshape(-1, 3)


def measure_center(pos, weight):
    pos, weight = f32(pos).reshape(-1), f32(weight).reshape(-1)
    com = np.zeros(3, np.float32)
    lib().vt_measure_center(_vp(pos), _vp(weight), weight.size, com.ctypes.data_as(c_float_p))
    return com


class FitContext:
    """resident pose-search state (target map, atoms, radii): vt_fit_create"""

    def __init__(self, pos, radii, target, resolution):
        _need()
        pos, radii = f32(pos).reshape(-1), f32(radii).reshape(-1)
        self.natoms = radii.size
        self.target = target  # borrowed by the library: keep it alive
        h = C.c_void_p()
        _check(lib().vt_fit_create(_vp(pos), _vp(radii), radii.size, target.h, resolution, C.byref(h)), "vt_fit_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) and _lib is not None and _inited:
            _lib.vt_fit_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        """what vt_fit_create chose (poses per batch, workspaces, list splat) + the host-preparation flag"""
        v = [C.c_int() for _ in range(4)]
        _check(lib().vt_fit_info(self.h, *[C.byref(x) for x in v]), "vt_fit_info")
        return {"batch": v[0].value, "nslots": v[1].value, "use_lists": bool(v[2].value), "prep_on_host": bool(v[3].value)}

    def set_origin(self, pos):
        pos = f32(pos).reshape(-1)
        _check(lib().vt_fit_set_origin(self.h, _vp(pos)), "vt_fit_set_origin")

    @staticmethod
    def make_poses(rot_deg, shift=None):
        rot = f32(rot_deg).reshape(-1, 3)
        sh = np.zeros_like(rot) if shift is None else f32(shift).reshape(-1, 3)
        return np.ascontiguousarray(np.concatenate([rot, sh], axis=1), np.float32)  # layout of vt_pose

    def eval_poses(self, com, poses):
        """poses: (n, 6) float32 = rot_deg xyz, shift xyz; returns float32 cc per pose (host buffers)"""
        poses = f32(poses).reshape(-1, 6)
        out = np.empty(len(poses), np.float32)
        cp, keep = _fp3(com)
        _check(lib().vt_fit_eval_poses(self.h, cp, _vp(poses), len(poses), _vp(out)), "vt_fit_eval_poses")
        return out

    def eval_poses_device(self, com, d_poses_ptr, nposes, d_cc_ptr):
        cp, keep = _fp3(com)
        _check(lib().vt_fit_eval_poses_device(self.h, cp, C.c_void_p(d_poses_ptr), nposes, C.c_void_p(d_cc_ptr)),
               "vt_fit_eval_poses_device")

    def search_grid(self, com, rot_deg, tsteps, tstep, want_all=False):
        """rotation x translation grid search (vt_fit_search_grid): (best_cc, best_index, best_pose(6,), cc_all)"""
        rot = f32(rot_deg).reshape(-1, 3)
        nt = 2 * tsteps + 1
        allcc = np.empty(len(rot) * nt ** 3, np.float32) if want_all else None
        cp, keep = _fp3(com)
        bcc, bidx, bpose = C.c_float(), C.c_long(), Pose()
        _check(lib().vt_fit_search_grid(self.h, cp, _vp(rot), len(rot), tsteps, tstep, C.byref(bcc), C.byref(bidx),
                                        C.byref(bpose), _vp(allcc) if want_all else None), "vt_fit_search_grid")
        pose = np.array(list(bpose.rot_deg) + list(bpose.shift), np.float32)
        return bcc.value, bidx.value, pose, allcc

    def rotate_table(self, com, stride, max_rot, first=0, step=1):
        tab = np.full(max_rot ** 3, np.nan, np.float32)
        cp, keep = _fp3(com)
        _check(lib().vt_fit_rotate_table(self.h, cp, stride, max_rot, first, step, _vp(tab)), "vt_fit_rotate_table")
        return tab


def fit(pos, radii, mass, target, resolution):
    """fit(), TclVoltool.C:679-928: returns (best coordinates, FitResult)"""
    _need()
    p, radii, mass = f32(pos).reshape(-1).copy(), f32(radii).reshape(-1), f32(mass).reshape(-1)
    res = FitResult()
    _check(lib().vt_fit(_vp(p), _vp(radii), _vp(mass), radii.size, target.h, resolution, C.byref(res)), "vt_fit")
    return p.reshape(-1, 3), res


def fit_schedule(pos, mass, dencom, evaluator):
    """the schedule alone with an injected evaluator (host logic; no CUDA touched).
    evaluator(origin_pos(n,3), com(3), stride, max_rot, first, step, table) fills table in place."""
    p, mass = f32(pos).reshape(-1).copy(), f32(mass).reshape(-1)
    n = mass.size

    def _cb(user, origin, com, stride, m, first, step, table):
        try:
            o = np.ctypeslib.as_array(origin, shape=(3 * n,)).reshape(-1, 3)
            c = np.ctypeslib.as_array(com, shape=(3,))
            t = np.ctypeslib.as_array(table, shape=(m * m * m,))
            evaluator(o, c, stride, m, first, step, t)
            return 0
        except Exception:  # surface as an error code; ctypes callbacks must not raise
            import traceback
            traceback.print_exc()
            return 9
    cb = EVAL_FN(_cb)
    res = FitResult()
    dc, keep = _fp3(dencom)
    _check(lib().vt_fit_schedule(_vp(p), _vp(mass), n, dc, cb, None, C.byref(res)), "vt_fit_schedule")
    return p.reshape(-1, 3), res
