// vt_api.cu -- the C ABI over the map kernels: cached-statistics bookkeeping and buffer
// replacement exactly as the VolumetricData methods do it (file:line in each function).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "vt_device.cuh"

using namespace vt;

namespace {

inline void inval_minmax(vt_cache &c) { c.minmax_valid = 0; c.cmin = 0; c.cmax = 0; }  // VolumetricData.C:1028-1032
inline void inval_mean(vt_cache &c) { c.mean_valid = 0; c.cmean = 0; }                  // :1089-1092
inline void inval_sigma(vt_cache &c) { c.sigma_valid = 0; c.csigma = 0; }               // :1103-1106

// compute_minmaxmean, VolumetricData.C:1047-1060 (minmaxmean_1fv_aligned: min, max and the mean
// as float(double sum / n))
int compute_minmaxmean(vt_map *m) {
  m->cache.cmin = 0; m->cache.cmax = 0; m->cache.cmean = 0;
  if (m->n > 0) {
    double s;
    int rc = k_minmaxsum(m->d, m->n, &m->cache.cmin, &m->cache.cmax, &s);
    if (rc) return rc;
    m->cache.cmean = (float)(s / m->n);
  }
  m->cache.mean_valid = 1; m->cache.minmax_valid = 1;
  return 0;
}
int compute_minmax(vt_map *m) {  // :1063-1074
  m->cache.cmin = 0; m->cache.cmax = 0;
  if (m->n > 0) {
    int rc = k_minmax(m->d, m->n, &m->cache.cmin, &m->cache.cmax);
    if (rc) return rc;
  }
  m->cache.minmax_valid = 1;
  return 0;
}
int compute_mean(vt_map *m) {  // :1109-1124
  double s = 0.0;
  if (m->n > 0) {
    int rc = k_sum(m->d, m->n, &s);
    if (rc) return rc;
  }
  s /= m->n;
  m->cache.cmean = (float)s;
  m->cache.mean_valid = 1;
  return 0;
}
int ensure_mean(vt_map *m) {  // ::mean :1078-1086
  if (!m->cache.mean_valid && !m->cache.minmax_valid) return compute_minmaxmean(m);
  if (!m->cache.mean_valid) return compute_mean(m);
  return 0;
}
int ensure_minmax(vt_map *m) {  // ::datarange :1035-1044
  if (!m->cache.minmax_valid && !m->cache.mean_valid) return compute_minmaxmean(m);
  if (!m->cache.minmax_valid) return compute_minmax(m);
  return 0;
}
int ensure_sigma(vt_map *m) {  // ::sigma / compute_sigma :1095-1100, 1127-1147
  if (m->cache.sigma_valid) return 0;
  int rc = ensure_mean(m);
  if (rc) return rc;
  double s = 0.0;
  if (m->n > 0) {
    rc = k_sumsq_delta(m->d, m->n, m->cache.cmean, &s);
    if (rc) return rc;
  }
  s /= m->n;
  s = std::sqrt(s);
  m->cache.csigma = (float)s;
  m->cache.sigma_valid = 1;
  return 0;
}

// swap in a freshly computed buffer (the reference's "delete[] data; data = data_new")
int replace_data(vt_map *m, float *fresh, const vt_grid &g) {
  int rc = dev_free(m->d);
  m->d = fresh;
  m->grid = g;
  m->n = gridsize(g);
  inval_minmax(m->cache); inval_mean(m->cache); inval_sigma(m->cache);
  return rc;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------ unary
int vt_clamp(vt_map *m, float lo, float hi) {
  VT_REQUIRE_CTX();
  int rc = k_unary(U_CLAMP, m->d, m->n, lo, hi, 0.f, 0.0);
  if (rc) return rc;
  if (m->cache.cmin < lo) m->cache.cmin = lo;  // :642-646
  if (m->cache.cmax > hi) m->cache.cmax = hi;
  inval_mean(m->cache); inval_sigma(m->cache);
  return 0;
}

int vt_scale_by(vt_map *m, float ff) {
  VT_REQUIRE_CTX();
  int rc = k_unary(U_SCALE, m->d, m->n, ff, 0.f, 0.f, 0.0);
  if (rc) return rc;
  vt_cache &c = m->cache;
  if (ff >= 0.0) { c.cmin *= ff; c.cmax *= ff; }                      // :662-670
  else { const float t = c.cmin; c.cmin = c.cmax * ff; c.cmax = t * ff; }
  inval_mean(c); inval_sigma(c);
  return 0;
}

int vt_scalar_add(vt_map *m, float ff) {
  VT_REQUIRE_CTX();
  int rc = k_unary(U_ADD, m->d, m->n, ff, 0.f, 0.f, 0.0);
  if (rc) return rc;
  m->cache.cmin += ff; m->cache.cmax += ff;                           // :686-688
  inval_mean(m->cache); inval_sigma(m->cache);
  return 0;
}

int vt_rescale_voxel_value_range(vt_map *m, float lo, float hi) {
  VT_REQUIRE_CTX();
  vt_cache &c = m->cache;
  const float newrange = hi - lo;                                     // :698-702 (reads cached_min/max as they are)
  const float oldrange = c.cmax - c.cmin;
  const float ratio = newrange / oldrange;
  int rc = k_unary(U_RESCALE, m->d, m->n, lo, ratio, c.cmin, 0.0);
  if (rc) return rc;
  c.cmin = lo; c.cmax = hi;
  inval_mean(c); inval_sigma(c);
  return 0;
}

int vt_sigma_scale(vt_map *m) {
  VT_REQUIRE_CTX();
  int rc = ensure_mean(m);
  if (rc) return rc;
  const float oldmean = m->cache.cmean;                               // :938-940
  rc = ensure_sigma(m);
  if (rc) return rc;
  const float inv = 1.0f / m->cache.csigma;
  rc = k_unary(U_SIGMA, m->d, m->n, oldmean, inv, 0.f, 0.0);
  if (rc) return rc;
  inval_minmax(m->cache); inval_mean(m->cache); inval_sigma(m->cache);
  return 0;
}

int vt_binmask(vt_map *m, float thr) {
  VT_REQUIRE_CTX();
  int rc = k_unary(U_BINMASK, m->d, m->n, thr, 0.f, 0.f, 0.0);
  if (rc) return rc;
  inval_minmax(m->cache); inval_mean(m->cache); inval_sigma(m->cache);
  return 0;
}

int vt_mdff_potential(vt_map *m, double threshold) {
  VT_REQUIRE_CTX();
  vt_cache &c = m->cache;
  const float threshinvert = (float)(threshold * -1);                 // :1001-1006
  const float mininvert = (c.cmax * -1);
  const float oldrange = threshinvert - mininvert;
  const float ratio = 1 / oldrange;
  int rc = k_unary(U_MDFFPOT, m->d, m->n, ratio, mininvert, 0.f, threshold);
  if (rc) return rc;
  c.cmin = 0; c.cmax = 1;
  inval_mean(c); inval_sigma(c);
  return 0;
}

// ------------------------------------------------------------------------------------------ stats
int vt_datarange(vt_map *m, float *mn, float *mx) {
  VT_REQUIRE_CTX();
  int rc = ensure_minmax(m);
  if (rc) return rc;
  *mn = m->cache.cmin; *mx = m->cache.cmax;
  return 0;
}
int vt_mean(vt_map *m, float *mean) {
  VT_REQUIRE_CTX();
  int rc = ensure_mean(m);
  if (rc) return rc;
  *mean = m->cache.cmean;
  return 0;
}
int vt_sigma(vt_map *m, float *sigma) {
  VT_REQUIRE_CTX();
  int rc = ensure_sigma(m);
  if (rc) return rc;
  *sigma = m->cache.csigma;
  return 0;
}

int vt_vol_com(const vt_map *m, float com[3]) {
  VT_REQUIRE_CTX();
  double acc[4];
  int rc = k_vol_com(m, acc);
  if (rc) return rc;
  const float scale = (float)(1.0 / acc[3]);                          // Voltool.C:243-245
  for (int k = 0; k < 3; ++k) com[k] = scale * (float)acc[k];
  return 0;
}

int vt_vol_moveto(vt_map *m, const float com[3], const float pos[3]) {
  vol_moveto(m->grid, com, pos);
  return 0;
}

int vt_vol_move(vt_map *m, const float mat[16]) {
  vol_move(m->grid, mat);
  return 0;
}

int vt_histogram(vt_map *m, int nbins, long *bins, float *midpts) {
  VT_REQUIRE_CTX();
  if (nbins < 1 || nbins > 8192) { set_error("vt_histogram: nbins out of range"); return 1; }
  float mn, mx;
  int rc = vt_datarange(m, &mn, &mx);
  if (rc) return rc;
  const double binwidth = (mx - mn) / nbins;                          // Voltool.C:447-449
  const double binwidthinv = 1 / binwidth;
  unsigned long long *d_bins;
  if (dev_alloc((void **)&d_bins, sizeof(unsigned long long) * nbins)) return 2;
  rc = k_histogram(m->d, m->n, mn, binwidthinv, nbins, d_bins);
  if (rc) { dev_free(d_bins); return rc; }
  unsigned long long *h = (unsigned long long *)malloc(sizeof(unsigned long long) * nbins);
  cudaError_t e = cudaMemcpyAsync(h, d_bins, sizeof(unsigned long long) * nbins, cudaMemcpyDeviceToHost, ctx().stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().stream);
  if (e != cudaSuccess) { free(h); dev_free(d_bins); return cuda_fail(e, "vt_histogram readback", __FILE__, __LINE__); }
  for (int j = 0; j < nbins; ++j) {
    bins[j] = (long)h[j];
    midpts[j] = (float)(mn + (0.5 * binwidth) + (j * binwidth));      // Voltool.C:458
  }
  free(h);
  dev_free(d_bins);
  return 0;
}

// ------------------------------------------------------------------------------------------ resizing
int vt_pad(vt_map *m, int pxm, int pxp, int pym, int pyp, int pzm, int pzp) {
  VT_REQUIRE_CTX();
  vt_grid g = m->grid;
  const int ox = g.xsize, oy = g.ysize;
  PadPlan p;
  pad_plan(g, pxm, pxp, pym, pyp, pzm, pzp, p);
  float *fresh;
  if (dev_alloc((void **)&fresh, sizeof(float) * (size_t)gridsize(g))) return 2;
  int rc = k_pad_copy(m->d, ox, oy, fresh, p);
  if (rc) { dev_free(fresh); return rc; }
  return replace_data(m, fresh, g);
}

int vt_crop(vt_map *m, double minx, double miny, double minz, double maxx, double maxy, double maxz) {
  int pads[6];
  crop_pads(m->grid, minx, miny, minz, maxx, maxy, maxz, pads);
  return vt_pad(m, pads[0], pads[1], pads[2], pads[3], pads[4], pads[5]);
}

int vt_downsample(vt_map *m) {
  VT_REQUIRE_CTX();
  vt_grid g = m->grid;
  const int ox = g.xsize, oy = g.ysize, oz = g.zsize;
  downsample_geom(g);
  float *fresh;
  if (dev_alloc((void **)&fresh, sizeof(float) * (size_t)gridsize(g))) return 2;
  int rc = k_downsample(m->d, ox, oy, oz, fresh);
  if (rc) return rc;
  return replace_data(m, fresh, g);
}

int vt_supersample(vt_map *m) {
  VT_REQUIRE_CTX();
  vt_grid g = m->grid;
  const int ox = g.xsize, oy = g.ysize, oz = g.zsize;
  supersample_geom(g);
  float *fresh;
  if (dev_alloc((void **)&fresh, sizeof(float) * (size_t)gridsize(g))) return 2;
  int rc = k_supersample(m->d, ox, oy, oz, fresh);
  if (rc) { dev_free(fresh); return rc; }
  return replace_data(m, fresh, g);
}

int vt_gaussian_blur(vt_map *m, double sigma) {
  VT_REQUIRE_CTX();
  float w[2 * 96 + 1];
  const int R = sigma > 0.0 ? blur_taps(sigma, w, 2 * 96 + 1) : -2;
  if (R == -1) { set_error("vt_gaussian_blur: sigma %g too large", sigma); return 1; }
  if (R >= 0 && m->n > 0) {
    float *fresh, *tmp;
    if (dev_alloc((void **)&fresh, sizeof(float) * (size_t)m->n)) return 2;
    if (dev_alloc((void **)&tmp, sizeof(float) * (size_t)m->n)) { dev_free(fresh); return 2; }
    int rc = k_blur(m->d, fresh, tmp, m->grid.xsize, m->grid.ysize, m->grid.zsize, w, R);
    dev_free(tmp);
    if (rc) { dev_free(fresh); return rc; }
    return replace_data(m, fresh, m->grid);
  }
  // sigma <= 0: data unchanged, caches still dropped (VolumetricData.C:990-996)
  inval_minmax(m->cache); inval_mean(m->cache); inval_sigma(m->cache);
  return 0;
}

// ------------------------------------------------------------------------------------------ cc / binary
int vt_cc(const vt_map *A, const vt_map *B, double threshold, double *cc, long *nused, double sums[5]) {
  VT_REQUIRE_CTX();
  double s[5];
  long n = 0;
  int rc = k_cc(A, B, threshold, s, &n);
  if (rc) return rc;
  *cc = cc_from_sums(s, n);
  if (nused) *nused = n;
  if (sums) memcpy(sums, s, sizeof(s));
  return 0;
}

int vt_cc_host(const vt_grid *gA, const float *a, const vt_grid *gB, const float *b, double threshold, double *cc) {
  VT_REQUIRE_CTX();
  vt_map *A = nullptr, *B = nullptr;
  int rc = vt_map_create(gA, a, &A);
  if (!rc) rc = vt_map_create(gB, b, &B);
  if (!rc) rc = vt_cc(A, B, threshold, cc, nullptr, nullptr);
  vt_map_destroy(A);
  vt_map_destroy(B);
  return rc;
}

int vt_init_from_intersection(const vt_grid *A, const vt_grid *B, vt_grid *out) { intersection_grid(*A, *B, *out); return 0; }
int vt_init_from_union(const vt_grid *A, const vt_grid *B, vt_grid *out) { union_grid(*A, *B, *out); return 0; }

int vt_binary(int op, int interp, int use_union, const vt_map *A, const vt_map *B, vt_map **out) {
  VT_REQUIRE_CTX();
  vt_grid og;
  if (use_union) union_grid(A->grid, B->grid, og);
  else intersection_grid(A->grid, B->grid, og);
  if (og.xsize < 0 || og.ysize < 0 || og.zsize < 0) { set_error("vt_binary: empty output grid"); return 1; }
  vt_map *o = new vt_map();
  o->grid = og;
  o->n = gridsize(og);
  memset(&o->cache, 0, sizeof(o->cache));
  if (dev_alloc((void **)&o->d, sizeof(float) * (size_t)o->n)) { delete o; return 2; }
  int rc = k_binary(op, interp, use_union, A, B, og, o->d);
  if (rc) { vt_map_destroy(o); return rc; }
  // init_new_volume_molecule -> VolumetricData ctor -> compute_minmax (Voltool.C:216-226)
  o->cache.cmin = 0; o->cache.cmax = 0;
  if (o->n > 0) rc = k_minmax(o->d, o->n, &o->cache.cmin, &o->cache.cmax);
  o->cache.minmax_valid = 1;
  *out = o;
  return rc;
}

int vt_binary_host(int op, int interp, int use_union, const vt_grid *gA, const float *a, const vt_grid *gB,
                   const float *b, vt_grid *out_grid, float **out_data) {
  VT_REQUIRE_CTX();
  vt_map *A = nullptr, *B = nullptr, *O = nullptr;
  int rc = vt_map_create(gA, a, &A);
  if (!rc) rc = vt_map_create(gB, b, &B);
  if (!rc) rc = vt_binary(op, interp, use_union, A, B, &O);
  if (!rc) {
    *out_grid = O->grid;
    *out_data = (float *)malloc(sizeof(float) * (size_t)(O->n > 0 ? O->n : 1));
    rc = vt_map_download(O, *out_data);
  }
  vt_map_destroy(A);
  vt_map_destroy(B);
  vt_map_destroy(O);
  return rc;
}

}  // extern "C"
