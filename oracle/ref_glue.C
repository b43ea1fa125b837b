/*
 * ref_glue.C -- links the reference's OWN object code (VolumetricData.C, Voltool.C, MDFF.C,
 * compiled unchanged from /root/reference by oracle/Makefile) into oracle/_ref/libvoltool_ref.so
 * and exposes it through a flat C API for ctypes.  TEST INFRASTRUCTURE ONLY.
 *
 * What is the reference's code here: cc_threaded, every VolumetricData method, every Voltool.C
 * function.  What is NOT (VMD sources missing from the snapshot, provided by the oracle's
 * documented restatements): Matrix4, GaussianBlur<float>, minmax*_1fv_aligned, stringdup,
 * WKF threads (a pthread tile scheduler), VMDApp (dummies).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define private public /* glue needs VolumetricData's cached-statistics members */
#include "VolumetricData.h"
#undef private
#include "Matrix4.h"
#include "Segmentation.h"
#include "VMDApp.h"
#include "MDFF.h"
#include "Voltool.h"
#include "voltool_oracle.h"

/* ---------------- link-time stand-ins ---------------- */
char *stringdup(const char *s) {
  char *r = new char[strlen(s) + 1];
  strcpy(r, s);
  return r;
}
void minmax_1fv_aligned(const float *f, long n, float *mn, float *mx) { vo_minmax_1fv(f, n, mn, mx); }
void minmaxmean_1fv_aligned(const float *f, long n, float *mn, float *mx, float *mean) {
  vo_minmaxmean_1fv(f, n, mn, mx, mean);
}

Matrix4::Matrix4() { vo_mat4_identity(mat); }
Matrix4::Matrix4(const float *m) { memcpy(mat, m, sizeof(mat)); }
int Matrix4::inverse() { return vo_mat4_inverse(mat); }
void Matrix4::multpoint3d(const float in[3], float out[3]) const { vo_mat4_multpoint3d(mat, in, out); }
void Matrix4::rotate_axis(const float axis[3], float angle) {
  int ax = (axis[0] != 0.f) ? 0 : ((axis[1] != 0.f) ? 1 : 2);
  vo_mat4_rot(mat, ax, angle);
}

template <> GaussianBlur<float>::GaussianBlur(float *img, int w, int h, int d, bool) : src(img), res(0), nx(w), ny(h), nz(d) {}
template <> void GaussianBlur<float>::blur(float sigma) {
  vo_grid g;
  memset(&g, 0, sizeof(g));
  g.xsize = nx; g.ysize = ny; g.zsize = nz;
  float *tmp = 0;
  vo_gaussian_blur(&g, src, &tmp, (double)sigma);
  long n = (long)nx * ny * nz;
  res = new float[n];
  memcpy(res, tmp, sizeof(float) * n);
  vo_free(tmp);
}
template <> float *GaussianBlur<float>::get_image() { return res; }

int VMDApp::molecule_new(const char *, int, int) { return 0; }
int VMDApp::molecule_add_volumetric(int, const char *, const double *, const double *, const double *, const double *,
                                    int, int, int, float *) { return 0; }
int VMDApp::molecule_set_style(const char *) { return 0; }
int VMDApp::molecule_addrep(int) { return 0; }

/* ---------------- WKF threads: pthread dynamic tile scheduler ---------------- */
static int g_numprocs = 0;
struct wkf_launch_state {
  pthread_mutex_t mtx;
  int cur, end;
  void *clientdata;
};
extern "C" int wkf_thread_numprocessors(void) {
  if (g_numprocs > 0) return g_numprocs;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
extern "C" int wkf_mutex_init(wkf_mutex_t *m) { return pthread_mutex_init(m, 0); }
extern "C" int wkf_mutex_lock(wkf_mutex_t *m) { return pthread_mutex_lock(m); }
extern "C" int wkf_mutex_unlock(wkf_mutex_t *m) { return pthread_mutex_unlock(m); }
extern "C" int wkf_mutex_destroy(wkf_mutex_t *m) { return pthread_mutex_destroy(m); }
extern "C" int wkf_threadlaunch_getdata(void *p, void **clientdata) {
  *clientdata = ((wkf_launch_state *)p)->clientdata;
  return 0;
}
extern "C" int wkf_threadlaunch_next_tile(void *p, int reqsize, wkf_tasktile_t *tile) {
  wkf_launch_state *s = (wkf_launch_state *)p;
  int rc = WKF_SCHED_DONE;
  pthread_mutex_lock(&s->mtx);
  if (s->cur < s->end) {
    tile->start = s->cur;
    tile->end = (s->end - s->cur > reqsize) ? s->cur + reqsize : s->end;
    s->cur = tile->end;
    rc = WKF_SCHED_CONTINUE;
  }
  pthread_mutex_unlock(&s->mtx);
  return rc;
}
extern "C" int wkf_threadlaunch(int numprocs, void *clientdata, void *fctn(void *), wkf_tasktile_t *tile) {
  wkf_launch_state s;
  pthread_mutex_init(&s.mtx, 0);
  s.cur = tile->start;
  s.end = tile->end;
  s.clientdata = clientdata;
  if (numprocs <= 1) {
    fctn(&s);
  } else {
    pthread_t *th = new pthread_t[numprocs];
    for (int i = 0; i < numprocs; i++) pthread_create(&th[i], 0, fctn, &s);
    for (int i = 0; i < numprocs; i++) pthread_join(th[i], 0);
    delete[] th;
  }
  pthread_mutex_destroy(&s.mtx);
  return 0;
}

/* ---------------- flat C API ---------------- */
static VolumetricData *make_vol(const vo_grid *g, const float *d) {
  long n = (long)g->xsize * g->ysize * g->zsize;
  float *copy = new float[n > 0 ? n : 1];
  if (d && n > 0) memcpy(copy, d, sizeof(float) * n);
  return new VolumetricData("ref", g->origin, g->xaxis, g->yaxis, g->zaxis, g->xsize, g->ysize, g->zsize, copy);
}
static void get_grid(const VolumetricData *v, vo_grid *g) {
  for (int k = 0; k < 3; k++) {
    g->origin[k] = v->origin[k]; g->xaxis[k] = v->xaxis[k]; g->yaxis[k] = v->yaxis[k]; g->zaxis[k] = v->zaxis[k];
  }
  g->xsize = v->xsize; g->ysize = v->ysize; g->zsize = v->zsize;
}

extern "C" {

void ref_set_numprocs(int n) { g_numprocs = n; }

void *ref_vol_new(const vo_grid *g, const float *d) { return make_vol(g, d); }
void ref_vol_free(void *h) { delete (VolumetricData *)h; }
void ref_vol_grid(void *h, vo_grid *g) { get_grid((VolumetricData *)h, g); }
const float *ref_vol_data(void *h) { return ((VolumetricData *)h)->data; }
void ref_vol_cache(void *h, vo_cache *c) {
  VolumetricData *v = (VolumetricData *)h;
  c->minmax_valid = v->minmax_isvalid; c->mean_valid = v->mean_isvalid; c->sigma_valid = v->sigma_isvalid;
  c->cmin = v->cached_min; c->cmax = v->cached_max; c->cmean = v->cached_mean; c->csigma = v->cached_sigma;
}

/* op codes shared with tests/refapi.py */
enum { R_CLAMP = 0, R_SCALE, R_SADD, R_RANGE, R_SIGMA, R_BINMASK, R_POT, R_DOWN, R_SUPER, R_BLUR, R_PAD, R_CROP };
void ref_vol_op(void *h, int op, const double *p) {
  VolumetricData *v = (VolumetricData *)h;
  switch (op) {
    case R_CLAMP: v->clamp((float)p[0], (float)p[1]); break;
    case R_SCALE: v->scale_by((float)p[0]); break;
    case R_SADD: v->scalar_add((float)p[0]); break;
    case R_RANGE: v->rescale_voxel_value_range((float)p[0], (float)p[1]); break;
    case R_SIGMA: v->sigma_scale(); break;
    case R_BINMASK: v->binmask((float)p[0]); break;
    case R_POT: v->mdff_potential(p[0]); break;
    case R_DOWN: v->downsample(); break;
    case R_SUPER: v->supersample(); break;
    case R_BLUR: v->gaussian_blur(p[0]); break;
    case R_PAD: v->pad((int)p[0], (int)p[1], (int)p[2], (int)p[3], (int)p[4], (int)p[5]); break;
    case R_CROP: v->crop(p[0], p[1], p[2], p[3], p[4], p[5]); break;
  }
}
void ref_vol_datarange(void *h, float *mn, float *mx) { ((VolumetricData *)h)->datarange(*mn, *mx); }
float ref_vol_mean(void *h) { return ((VolumetricData *)h)->mean(); }
float ref_vol_sigma(void *h) { return ((VolumetricData *)h)->sigma(); }
void ref_vol_com(void *h, float *com) { vol_com((VolumetricData *)h, com); }
void ref_vol_moveto(void *h, float *com, float *pos) { vol_moveto((VolumetricData *)h, com, pos); }
void ref_vol_move(void *h, float *mat) { vol_move((VolumetricData *)h, mat); }
void ref_vol_histogram(void *h, int nbins, long *bins, float *midpts) {
  /* one spare slot: the reference writes bins[nbins] for the maximum voxel (Voltool.C:457) */
  long *tmp = new long[nbins + 1];
  tmp[nbins] = 0;
  histogram((VolumetricData *)h, nbins, tmp, midpts);
  memcpy(bins, tmp, sizeof(long) * nbins);
  delete[] tmp;
}

void ref_voxel_coord(void *h, long i, float *xyz) { ((VolumetricData *)h)->voxel_coord(i, xyz[0], xyz[1], xyz[2]); }
void ref_voxel_coord_int(void *h, int i, float *xyz) { voxel_coord(i, xyz[0], xyz[1], xyz[2], (VolumetricData *)h); }
void ref_voxel_coord_from_cartesian(void *h, const float *car, float *vox, int shiftflag) {
  ((VolumetricData *)h)->voxel_coord_from_cartesian_coord(car, vox, shiftflag);
}
long ref_voxel_index_from_coord(void *h, float x, float y, float z) {
  return ((VolumetricData *)h)->voxel_index_from_coord(x, y, z);
}
/* mode: 0 nearest, 1 nearest safe, 2 interp, 3 interp safe */
void ref_lookup(void *h, int mode, long n, const float *xyz, float *out) {
  VolumetricData *v = (VolumetricData *)h;
  for (long i = 0; i < n; i++) {
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    switch (mode) {
      case 0: out[i] = v->voxel_value_from_coord(x, y, z); break;
      case 1: out[i] = v->voxel_value_from_coord_safe(x, y, z); break;
      case 2: out[i] = v->voxel_value_interpolate_from_coord(x, y, z); break;
      default: out[i] = v->voxel_value_interpolate_from_coord_safe(x, y, z); break;
    }
  }
}

void ref_init_from_intersection(const vo_grid *A, const vo_grid *B, vo_grid *o) {
  VolumetricData *a = make_vol(A, 0), *b = make_vol(B, 0);
  VolumetricData *nv = init_new_volume();
  init_from_intersection(a, (const VolumetricData *)b, nv);
  get_grid(nv, o);
  delete a; delete b; delete nv;
}
void ref_init_from_union(const vo_grid *A, const vo_grid *B, vo_grid *o) {
  VolumetricData *a = make_vol(A, 0), *b = make_vol(B, 0);
  VolumetricData *nv = init_new_volume();
  init_from_union(a, (const VolumetricData *)b, nv);
  get_grid(nv, o);
  delete a; delete b; delete nv;
}

int ref_cc_vols(void *a, void *b, double thr, double *cc) {
  return cc_threaded((VolumetricData *)a, (const VolumetricData *)b, cc, thr);
}
int ref_cc(const vo_grid *A, const float *a, const vo_grid *B, const float *b, double thr, double *cc) {
  VolumetricData *va = make_vol(A, a), *vb = make_vol(B, b);
  int rc = cc_threaded(va, vb, cc, thr);
  delete va; delete vb;
  return rc;
}

/* op: 0 add 1 subtract 2 multiply 3 average.  *out is malloc()ed (free with vo_free). */
void ref_binary(int op, int interp, int use_union, const vo_grid *A, const float *a, const vo_grid *B, const float *b,
                vo_grid *og, float **out) {
  VolumetricData *va = make_vol(A, a), *vb = make_vol(B, b);
  VolumetricData *nv = init_new_volume();
  switch (op) {
    case 0: add(va, vb, nv, interp != 0, use_union != 0); break;
    case 1: subtract(va, vb, nv, interp != 0, use_union != 0); break;
    case 2: multiply(va, vb, nv, interp != 0, use_union != 0); break;
    default: average(va, vb, nv, interp != 0, use_union != 0); break;
  }
  get_grid(nv, og);
  long n = nv->gridsize();
  float *o = (float *)malloc(sizeof(float) * (n > 0 ? n : 1));
  if (n > 0) memcpy(o, nv->data, sizeof(float) * n);
  *out = o;
  delete va; delete vb; delete nv;
}

/* cc backend for the oracle's restated fit driver: the reference's cc_threaded on a persistent
   target volume (ctx = handle from ref_vol_new) */
static double ref_cc_backend(const vo_grid *A, const float *a, const vo_grid *B, const float *b, double thr, void *ctx) {
  (void)B; (void)b;
  VolumetricData *va = make_vol(A, a);
  double cc = 0.0;
  cc_threaded(va, (const VolumetricData *)ctx, &cc, thr);
  delete va;
  return cc;
}
void ref_install_cc_backend(void *target_handle) {
  if (target_handle) vo_set_cc_backend(ref_cc_backend, target_handle);
  else vo_set_cc_backend(0, 0);
}

} /* extern "C" */
