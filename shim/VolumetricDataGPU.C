/*
 * VolumetricDataGPU.C -- the reference-side adapter: the numeric-core entry points of voltool with bodies that
 * forward to libvoltool_gpu.so (include/voltool_gpu.h).  Compiled against the reference's OWN headers
 * (VolumetricData.h, Voltool.h, MDFF.h); a VMD maintainer drops this file into the build in place of the bodies at
 *
 *     MDFF.C:102-151               cc_threaded
 *     VolumetricData.C:542-631     pad, crop
 *     VolumetricData.C:634-714     clamp, scale_by, scalar_add, rescale_voxel_value_range
 *     VolumetricData.C:720-932     downsample, supersample
 *     VolumetricData.C:937-1025    sigma_scale, binmask, gaussian_blur, mdff_potential
 *     VolumetricData.C:1035-1147   datarange, mean, sigma
 *     Voltool.C:229-247            vol_com
 *     Voltool.C:249-377            add, subtract, multiply, average
 *     Voltool.C:443-461            histogram
 *
 * and keeps everything else of those three files (constructors, destructor, cell_axes, the voxel lookups, the
 * gradient code, init_from_*, vol_move / vol_moveto ...).  shim/Makefile builds exactly that combination here: the
 * reference's object files with their symbols weakened (objcopy --weaken) linked with this file, so that the bodies
 * below win and every other symbol is still the reference's; tests/test_gpu_shim.py then drives the reference's
 * VolumetricData class through both libraries (oracle/_ref = all CPU, shim/_build = these bodies) and compares.
 *
 * Model: "one call = one reference call".  VolumetricData keeps owning `data` on the host (VolumetricData.C:95-99);
 * a call uploads the voxels together with the object's cached-statistics block, runs the resident-map entry point,
 * and writes voxels, geometry and caches back -- the library maintains the caches by the reference's own rules, so
 * after the call the object is in the state the CPU body would have left it in.  Chains of commands that should
 * stay in HBM use the vt_map handles directly.
 * Errors: the reference's bodies are void; a failing GPU call prints vt_last_error() and leaves the object
 * untouched (cc_threaded, the one int-returning entry point, returns the status, like the old CUDA hooks
 * TclMDFF.C:140-152).  There is no CPU fallback.
 */
#include <stdio.h>
#include <string.h>

#define private public /* the bodies below are VolumetricData members in the maintainer's build */
#include "VolumetricData.h"
#undef private
#include "MDFF.h"
#include "VMDApp.h"
#include "Voltool.h"
#include "voltool_gpu.h"

namespace {

vt_grid grid_of(const VolumetricData *v) {
  vt_grid g;
  for (int i = 0; i < 3; i++) {
    g.origin[i] = v->origin[i]; g.xaxis[i] = v->xaxis[i]; g.yaxis[i] = v->yaxis[i]; g.zaxis[i] = v->zaxis[i];
  }
  g.xsize = v->xsize; g.ysize = v->ysize; g.zsize = v->zsize;
  return g;
}

vt_cache cache_of(const VolumetricData *v) {
  vt_cache c;
  c.minmax_valid = v->minmax_isvalid; c.mean_valid = v->mean_isvalid; c.sigma_valid = v->sigma_isvalid;
  c.cmin = v->cached_min; c.cmax = v->cached_max; c.cmean = v->cached_mean; c.csigma = v->cached_sigma;
  return c;
}

int fail(const char *where) {
  fprintf(stderr, "voltool_gpu: %s: %s\n", where, vt_last_error());
  return 1;
}

/* the resident twin of v as it stands (voxels + caches) */
vt_map *resident(const VolumetricData *v) {
  vt_grid g = grid_of(v);
  vt_cache c = cache_of(v);
  vt_map *m = NULL;
  static const float none = 0.0f;
  if (vt_map_adopt_host(&g, v->data ? v->data : &none, &c, &m)) { fail("upload"); return NULL; }
  return m;
}

/* caches only (the statistics calls) */
void take_cache(VolumetricData *v, const vt_map *m) {
  vt_cache c;
  vt_map_cache(m, &c);
  v->minmax_isvalid = c.minmax_valid != 0; v->mean_isvalid = c.mean_valid != 0; v->sigma_isvalid = c.sigma_valid != 0;
  v->cached_min = c.cmin; v->cached_max = c.cmax; v->cached_mean = c.cmean; v->cached_sigma = c.csigma;
}

/* voxels, geometry and caches back into v; the buffer is replaced when the op resized the map ("delete[] data;
   data = data_new", e.g. VolumetricData.C:604-606); every mutating body but scalar_add ends in invalidate_gradient() */
void writeback(VolumetricData *v, vt_map *m, bool drop_gradient) {
  vt_grid g;
  vt_map_grid(m, &g);
  const long n = vt_map_gridsize(m);
  if (n != v->gridsize() || !v->data) {
    float *fresh = new float[n > 0 ? n : 1];
    if (vt_map_download(m, fresh)) { fail("download"); delete[] fresh; vt_map_destroy(m); return; }
    delete[] v->data;
    v->data = fresh;
  } else if (vt_map_download(m, v->data)) {
    fail("download");
    vt_map_destroy(m);
    return;
  }
  for (int i = 0; i < 3; i++) {
    v->origin[i] = g.origin[i]; v->xaxis[i] = g.xaxis[i]; v->yaxis[i] = g.yaxis[i]; v->zaxis[i] = g.zaxis[i];
  }
  v->xsize = g.xsize; v->ysize = g.ysize; v->zsize = g.zsize;
  take_cache(v, m);
  if (drop_gradient) v->invalidate_gradient();
  vt_map_destroy(m);
}

}  // namespace

#define VT_MUTATE_G(call, what, drop_gradient)     \
  do {                                             \
    vt_map *m = resident(this);                    \
    if (!m) return;                                \
    if (call) { fail(what); vt_map_destroy(m); return; } \
    writeback(this, m, drop_gradient);             \
  } while (0)
#define VT_MUTATE(call, what) VT_MUTATE_G(call, what, true)

/* ---------------------------------------------------------------- VolumetricData.C */
void VolumetricData::pad(int padxm, int padxp, int padym, int padyp, int padzm, int padzp) {
  VT_MUTATE(vt_pad(m, padxm, padxp, padym, padyp, padzm, padzp), "pad");
}
void VolumetricData::crop(double crop_minx, double crop_miny, double crop_minz, double crop_maxx, double crop_maxy,
                          double crop_maxz) {
  VT_MUTATE(vt_crop(m, crop_minx, crop_miny, crop_minz, crop_maxx, crop_maxy, crop_maxz), "crop");
}
void VolumetricData::clamp(float min_value, float max_value) { VT_MUTATE(vt_clamp(m, min_value, max_value), "clamp"); }
void VolumetricData::scale_by(float ff) { VT_MUTATE(vt_scale_by(m, ff), "scale_by"); }
void VolumetricData::scalar_add(float ff) {
  /* the one mutating body that leaves the gradient in place (VolumetricData.C:680-694) */
  VT_MUTATE_G(vt_scalar_add(m, ff), "scalar_add", false);
}
void VolumetricData::rescale_voxel_value_range(float min_value, float max_value) {
  VT_MUTATE(vt_rescale_voxel_value_range(m, min_value, max_value), "rescale_voxel_value_range");
}
void VolumetricData::downsample() { VT_MUTATE(vt_downsample(m), "downsample"); }
void VolumetricData::supersample() { VT_MUTATE(vt_supersample(m), "supersample"); }
void VolumetricData::sigma_scale() { VT_MUTATE(vt_sigma_scale(m), "sigma_scale"); }
void VolumetricData::binmask(float threshold) { VT_MUTATE(vt_binmask(m, threshold), "binmask"); }
void VolumetricData::gaussian_blur(double sigma) { VT_MUTATE(vt_gaussian_blur(m, sigma), "gaussian_blur"); }
void VolumetricData::mdff_potential(double threshold) { VT_MUTATE(vt_mdff_potential(m, threshold), "mdff_potential"); }

void VolumetricData::datarange(float &min, float &max) {
  vt_map *m = resident(this);
  if (!m) return;
  if (vt_datarange(m, &min, &max)) fail("datarange");
  else take_cache(this, m);
  vt_map_destroy(m);
}
float VolumetricData::mean() {
  float r = 0.0f;
  vt_map *m = resident(this);
  if (!m) return r;
  if (vt_mean(m, &r)) fail("mean");
  else take_cache(this, m);
  vt_map_destroy(m);
  return r;
}
float VolumetricData::sigma() {
  float r = 0.0f;
  vt_map *m = resident(this);
  if (!m) return r;
  if (vt_sigma(m, &r)) fail("sigma");
  else take_cache(this, m);
  vt_map_destroy(m);
  return r;
}

/* ---------------------------------------------------------------- MDFF.C */
int cc_threaded(VolumetricData *qsVol, const VolumetricData *targetVol, double *cc, double threshold) {
  vt_grid ga = grid_of(qsVol), gb = grid_of(targetVol);
  if (vt_cc_host(&ga, qsVol->data, &gb, targetVol->data, threshold, cc)) return fail("cc_threaded");
  return 0;
}

/* ---------------------------------------------------------------- Voltool.C */
namespace {
void two_map_op(int op, VolumetricData *mapA, VolumetricData *mapB, VolumetricData *newvol, bool interp, bool USE_UNION) {
  vt_grid ga = grid_of(mapA), gb = grid_of(mapB), go;
  float *out = NULL;
  if (vt_binary_host(op, interp ? 1 : 0, USE_UNION ? 1 : 0, &ga, mapA->data, &gb, mapB->data, &go, &out)) {
    fail("two-map op");
    return;
  }
  /* init_from_union / init_from_intersection (Voltool.C:51-147): geometry + a fresh buffer; newvol's caches are
     left exactly as they were, as in the reference */
  for (int i = 0; i < 3; i++) {
    newvol->origin[i] = go.origin[i]; newvol->xaxis[i] = go.xaxis[i]; newvol->yaxis[i] = go.yaxis[i];
    newvol->zaxis[i] = go.zaxis[i];
  }
  newvol->xsize = go.xsize; newvol->ysize = go.ysize; newvol->zsize = go.zsize;
  if (newvol->data) delete[] newvol->data;
  const long n = newvol->gridsize();
  newvol->data = new float[n > 0 ? n : 1];
  if (n > 0) memcpy(newvol->data, out, sizeof(float) * n);
  vt_free_host(out);
}
}  // namespace

void add(VolumetricData *mapA, VolumetricData *mapB, VolumetricData *newvol, bool interp, bool USE_UNION) {
  two_map_op(VT_ADD, mapA, mapB, newvol, interp, USE_UNION);
}
void subtract(VolumetricData *mapA, VolumetricData *mapB, VolumetricData *newvol, bool interp, bool USE_UNION) {
  two_map_op(VT_SUBTRACT, mapA, mapB, newvol, interp, USE_UNION);
}
void multiply(VolumetricData *mapA, VolumetricData *mapB, VolumetricData *newvol, bool interp, bool USE_UNION) {
  two_map_op(VT_MULTIPLY, mapA, mapB, newvol, interp, USE_UNION);
}
void average(VolumetricData *mapA, VolumetricData *mapB, VolumetricData *newvol, bool interp, bool USE_UNION) {
  two_map_op(VT_AVERAGE, mapA, mapB, newvol, interp, USE_UNION);
}

void vol_com(VolumetricData *vol, float *com) {
  vt_map *m = resident(vol);
  if (!m) return;
  if (vt_vol_com(m, com)) fail("vol_com");
  vt_map_destroy(m);
}

void histogram(VolumetricData *vol, int nbins, long *bins, float *midpts) {
  vt_map *m = resident(vol);
  if (!m) return;
  if (vt_histogram(m, nbins, bins, midpts)) fail("histogram");
  else take_cache(vol, m);   /* histogram() calls datarange() first (Voltool.C:445) */
  vt_map_destroy(m);
}
