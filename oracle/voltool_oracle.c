/*
 * voltool_oracle.c -- CPU restatement of voltool's hot path.  TEST INFRASTRUCTURE ONLY.
 * See voltool_oracle.h for the rules (who may call this) and the parity-pin status.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off and no -ffast-math on purpose: every float/double operation below rounds
 * exactly where the C++ reference's expression does (FLT_EVAL_METHOD == 0 on x86-64), so
 * "bit-exact" is well defined.  The reference's own build uses -O6 -ffast-math (configure:2336).
 *
 * Citations are file:line into the reference snapshot.
 */
#include "voltool_oracle.h"

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VO_MIN(a, b) (((a) < (b)) ? (a) : (b))
#define VO_MAX(a, b) (((a) > (b)) ? (a) : (b))
#define VO_PI 3.14159265358979323846 /* utilities.h:41 */

static float vo_nan(void) {
  union { uint32_t u; float f; } v;
  v.u = 0x7fc00000u;
  return v.f;
}

/* Voltool.h:32-35 */
static int vo_isnan(float f) {
  union { float f; uint32_t x; } u;
  u.f = f;
  return (u.x << 1) > 0xff000000u;
}

/* utilities.h:141-143 */
static double dot3d(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

long vo_gridsize(const vo_grid *g) { return (long)g->xsize * (long)g->ysize * (long)g->zsize; }
void vo_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * Matrix4 [UNPINNED: Matrix4.C is not in the snapshot].  Column-major float[16], translation
 * in m[12..14] (as used at VolumetricData.C:207-223, Voltool.C:405-413).
 * ------------------------------------------------------------------------------------------ */
void vo_mat4_identity(float m[16]) {
  for (int i = 0; i < 16; i++) m[i] = 0.0f;
  m[0] = m[5] = m[10] = m[15] = 1.0f;
}

/* Gauss-Jordan on the augmented [M | I], all arithmetic in float, eliminating on the DIAGONAL
   pivots in order; rows are exchanged only when a diagonal pivot is exactly zero (then with the
   row below holding the largest entry of that column).  For the cell matrices voltool builds
   (diagonal 3x3 + translation row, VolumetricData.C:207-223) no exchange happens and every
   structural zero of the inverse stays an exact zero.  Returns 1 if singular. */
int vo_mat4_inverse(float m[16]) {
  float w[4][8];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      w[r][c] = m[4 * r + c];
      w[r][4 + c] = (r == c) ? 1.0f : 0.0f;
    }
  for (int i = 0; i < 4; i++) {
    if (w[i][i] == 0.0f) {
      int p = -1;
      float big = 0.0f;
      for (int j = i + 1; j < 4; j++)
        if (fabsf(w[j][i]) > big) { big = fabsf(w[j][i]); p = j; }
      if (p < 0) return 1;
      for (int c = 0; c < 8; c++) { float t = w[i][c]; w[i][c] = w[p][c]; w[p][c] = t; }
    }
    float s = 1.0f / w[i][i];
    for (int c = 0; c < 8; c++) w[i][c] *= s;
    for (int r = 0; r < 4; r++) {
      if (r == i) continue;
      float f = w[r][i];
      if (f == 0.0f) continue;
      for (int c = 0; c < 8; c++) w[r][c] -= w[i][c] * f;
    }
  }
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) m[4 * r + c] = w[r][4 + c];
  return 0;
}

/* projective point transform: divide by w first, then the affine part (used at
   VolumetricData.C:227 and by measure_move, TclVoltool.C:62) */
void vo_mat4_multpoint3d(const float m[16], const float in[3], float out[3]) {
  float iw = 1.0f / (in[0] * m[3] + in[1] * m[7] + in[2] * m[11] + m[15]);
  float t0 = iw * in[0], t1 = iw * in[1], t2 = iw * in[2];
  float o0 = t0 * m[0] + t1 * m[4] + t2 * m[8] + iw * m[12];
  float o1 = t0 * m[1] + t1 * m[5] + t2 * m[9] + iw * m[13];
  float o2 = t0 * m[2] + t1 * m[6] + t2 * m[10] + iw * m[14];
  out[0] = o0; out[1] = o1; out[2] = o2;
}

/* Matrix4::rotate_axis for the unit axes used at TclVoltool.C:167-174: a right-handed rotation
   by angle_rad about axis 0/1/2, cos/sin evaluated in double then stored as float. */
void vo_mat4_rot(float m[16], int axis, float angle_rad) {
  vo_mat4_identity(m);
  float c = (float)cos((double)angle_rad);
  float s = (float)sin((double)angle_rad);
  if (axis == 0) { m[5] = c; m[6] = s; m[9] = -s; m[10] = c; }
  else if (axis == 1) { m[0] = c; m[2] = -s; m[8] = s; m[10] = c; }
  else { m[0] = c; m[1] = s; m[4] = -s; m[5] = c; }
}

/* ------------------------------------------------------------------------------------------
 * Geometry and lookups [PINNED]
 * ------------------------------------------------------------------------------------------ */
/* VolumetricData.C:138-168 */
void vo_cell_axes(const vo_grid *g, float xax[3], float yax[3], float zax[3]) {
  float xsinv = (g->xsize > 1) ? 1.0f / (g->xsize - 1.0f) : 1.0f;
  float ysinv = (g->ysize > 1) ? 1.0f / (g->ysize - 1.0f) : 1.0f;
  float zsinv = (g->zsize > 1) ? 1.0f / (g->zsize - 1.0f) : 1.0f;
  for (int k = 0; k < 3; k++) {
    xax[k] = (float)(g->xaxis[k] * xsinv);
    yax[k] = (float)(g->yaxis[k] * ysinv);
    zax[k] = (float)(g->zaxis[k] * zsinv);
  }
}

/* VolumetricData.h:142-153 (member, long index) */
void vo_voxel_coord(const vo_grid *g, long i, float *x, float *y, float *z) {
  float xd[3], yd[3], zd[3];
  vo_cell_axes(g, xd, yd, zd);
  long gz = i / (g->ysize * g->xsize);
  long gy = (i / g->xsize) % g->ysize;
  long gx = i % g->xsize;
  *x = (float)(g->origin[0] + (gx * xd[0]) + (gy * yd[0]) + (gz * zd[0]));
  *y = (float)(g->origin[1] + (gx * xd[1]) + (gy * yd[1]) + (gz * zd[1]));
  *z = (float)(g->origin[2] + (gx * xd[2]) + (gy * yd[2]) + (gz * zd[2]));
}

/* Voltool.h:63-77 (free function, int index; used by cc) */
void vo_voxel_coord_int(const vo_grid *g, int i, float *x, float *y, float *z) {
  float xd[3], yd[3], zd[3];
  vo_cell_axes(g, xd, yd, zd);
  int xsize = g->xsize, ysize = g->ysize;
  int gz = i / (ysize * xsize);
  int gy = (i / xsize) % ysize;
  int gx = i % xsize;
  *x = (float)(g->origin[0] + (gx * xd[0]) + (gy * yd[0]) + (gz * zd[0]));
  *y = (float)(g->origin[1] + (gx * xd[1]) + (gy * yd[1]) + (gz * zd[1]));
  *z = (float)(g->origin[2] + (gx * xd[2]) + (gy * yd[2]) + (gz * zd[2]));
}

/* The loop-invariant part of VolumetricData.C:194-225: build the cell matrix and invert it. */
void vo_grid_inverse(const vo_grid *g, int shiftflag, float inv[16]) {
  float mc[3];
  int xs = g->xsize, ys = g->ysize, zs = g->zsize;
  if (shiftflag) {
    for (int i = 0; i < 3; i++)
      mc[i] = (float)g->origin[i] -
              0.5f * (float)(g->xaxis[i] / (xs - 1) + g->yaxis[i] / (ys - 1) + g->zaxis[i] / (zs - 1));
  } else {
    for (int i = 0; i < 3; i++) mc[i] = (float)g->origin[i];
  }
  inv[0] = (float)g->xaxis[0] / (xs - 1);
  inv[1] = (float)g->yaxis[0] / (ys - 1);
  inv[2] = (float)g->zaxis[0] / (zs - 1);
  inv[3] = 0.f;
  inv[4] = (float)g->xaxis[1] / (xs - 1);
  inv[5] = (float)g->yaxis[1] / (ys - 1);
  inv[6] = (float)g->zaxis[1] / (zs - 1);
  inv[7] = 0.f;
  inv[8] = (float)g->xaxis[2] / (xs - 1);
  inv[9] = (float)g->yaxis[2] / (ys - 1);
  inv[10] = (float)g->zaxis[2] / (zs - 1);
  inv[11] = 0.f;
  inv[12] = mc[0];
  inv[13] = mc[1];
  inv[14] = mc[2];
  inv[15] = 1.f;
  vo_mat4_inverse(inv);
}

/* VolumetricData.C:194-228 */
void vo_voxel_coord_from_cartesian(const vo_grid *g, const float car[3], float vox[3], int shiftflag) {
  float inv[16];
  vo_grid_inverse(g, shiftflag, inv);
  vo_mat4_multpoint3d(inv, car, vox);
}

/* VolumetricData.C:253-260 */
float vo_voxel_value_safe(const vo_grid *g, const float *d, int x, int y, int z) {
  long xx = (x > 0) ? ((x < g->xsize) ? x : g->xsize - 1) : 0;
  long yy = (y > 0) ? ((y < g->ysize) ? y : g->ysize - 1) : 0;
  long zz = (z > 0) ? ((z < g->zsize) ? z : g->zsize - 1) : 0;
  return d[zz * g->xsize * g->ysize + yy * g->xsize + xx];
}

/* VolumetricData.C:264-291 */
float vo_voxel_value_interpolate(const vo_grid *g, const float *d, float xv, float yv, float zv) {
  int x = (int)xv, y = (int)yv, z = (int)zv;
  float xf = xv - x, yf = yv - y, zf = zv - z;
  float xl[4], yl[2], tmp;
  tmp = vo_voxel_value_safe(g, d, x, y, z);
  xl[0] = tmp + xf * (vo_voxel_value_safe(g, d, x + 1, y, z) - tmp);
  tmp = vo_voxel_value_safe(g, d, x, y + 1, z);
  xl[1] = tmp + xf * (vo_voxel_value_safe(g, d, x + 1, y + 1, z) - tmp);
  tmp = vo_voxel_value_safe(g, d, x, y, z + 1);
  xl[2] = tmp + xf * (vo_voxel_value_safe(g, d, x + 1, y, z + 1) - tmp);
  tmp = vo_voxel_value_safe(g, d, x, y + 1, z + 1);
  xl[3] = tmp + xf * (vo_voxel_value_safe(g, d, x + 1, y + 1, z + 1) - tmp);
  yl[0] = xl[0] + yf * (xl[1] - xl[0]);
  yl[1] = xl[2] + yf * (xl[3] - xl[2]);
  return yl[0] + zf * (yl[1] - yl[0]);
}

/* with a precomputed inverse (identical values: the inverse is loop-invariant) */
static long index_from_coord_inv(const vo_grid *g, const float inv1[16], float x, float y, float z) {
  float car[3] = {x, y, z}, gp[3];
  vo_mat4_multpoint3d(inv1, car, gp);
  int gx = (int)gp[0], gy = (int)gp[1], gz = (int)gp[2];
  if (gx < 0 || gx >= g->xsize) return -1;
  if (gy < 0 || gy >= g->ysize) return -1;
  if (gz < 0 || gz >= g->zsize) return -1;
  return gz * (long)g->xsize * (long)g->ysize + gy * (long)g->xsize + gx;
}

/* VolumetricData.C:232-249 */
long vo_voxel_index_from_coord(const vo_grid *g, float x, float y, float z) {
  float inv[16];
  vo_grid_inverse(g, 1, inv);
  return index_from_coord_inv(g, inv, x, y, z);
}

/* VolumetricData.C:295-301 / 327-333 (note the ind > 0 test: voxel 0 reads as outside) */
static float value_from_coord_inv(const vo_grid *g, const float *d, const float inv1[16], float x, float y,
                                  float z, int safe) {
  long ind = index_from_coord_inv(g, inv1, x, y, z);
  if (ind > 0) return d[ind];
  return safe ? 0.0f : vo_nan();
}

float vo_value_from_coord(const vo_grid *g, const float *d, float x, float y, float z, int safe) {
  float inv[16];
  vo_grid_inverse(g, 1, inv);
  return value_from_coord_inv(g, d, inv, x, y, z, safe);
}

/* VolumetricData.C:305-322 / 338-355 */
static float interp_from_coord_inv(const vo_grid *g, const float *d, const float inv0[16], float x, float y,
                                   float z, int safe) {
  float car[3] = {x, y, z}, gp[3];
  vo_mat4_multpoint3d(inv0, car, gp);
  int gx = (int)gp[0], gy = (int)gp[1], gz = (int)gp[2];
  float outside = safe ? 0.0f : vo_nan();
  if (gx < 0 || gx >= g->xsize) return outside;
  if (gy < 0 || gy >= g->ysize) return outside;
  if (gz < 0 || gz >= g->zsize) return outside;
  return vo_voxel_value_interpolate(g, d, gp[0], gp[1], gp[2]);
}

float vo_value_interp_from_coord(const vo_grid *g, const float *d, float x, float y, float z, int safe) {
  float inv[16];
  vo_grid_inverse(g, 0, inv);
  return interp_from_coord_inv(g, d, inv, x, y, z, safe);
}

/* Voltool.C:51-78 (data allocation left to the caller) */
void vo_init_from_intersection(const vo_grid *A, const vo_grid *B, vo_grid *o) {
  for (int d = 0; d < 3; d++) {
    o->origin[d] = VO_MAX(A->origin[d], B->origin[d]);
    o->xaxis[d] = VO_MAX(VO_MIN(A->origin[d] + A->xaxis[d], B->origin[d] + B->xaxis[d]), o->origin[d]);
    o->yaxis[d] = VO_MAX(VO_MIN(A->origin[d] + A->yaxis[d], B->origin[d] + B->yaxis[d]), o->origin[d]);
    o->zaxis[d] = VO_MAX(VO_MIN(A->origin[d] + A->zaxis[d], B->origin[d] + B->zaxis[d]), o->origin[d]);
  }
  for (int d = 0; d < 3; d++) {
    o->xaxis[d] = o->xaxis[d] - o->origin[d];
    o->yaxis[d] = o->yaxis[d] - o->origin[d];
    o->zaxis[d] = o->zaxis[d] - o->origin[d];
  }
  o->xsize = (int)VO_MAX(dot3d(o->xaxis, A->xaxis) * A->xsize / dot3d(A->xaxis, A->xaxis),
                        dot3d(o->xaxis, B->xaxis) * B->xsize / dot3d(B->xaxis, B->xaxis));
  o->ysize = (int)VO_MAX(dot3d(o->yaxis, A->yaxis) * A->ysize / dot3d(A->yaxis, A->yaxis),
                        dot3d(o->yaxis, B->yaxis) * B->ysize / dot3d(B->yaxis, B->yaxis));
  o->zsize = (int)VO_MAX(dot3d(o->zaxis, A->zaxis) * A->zsize / dot3d(A->zaxis, A->zaxis),
                        dot3d(o->zaxis, B->zaxis) * B->zsize / dot3d(B->zaxis, B->zaxis));
}

/* Voltool.C:114-147 */
void vo_init_from_union(const vo_grid *A, const vo_grid *B, vo_grid *o) {
  for (int d = 0; d < 3; d++) { o->xaxis[d] = 0.0; o->yaxis[d] = 0.0; o->zaxis[d] = 0.0; }
  for (int d = 0; d < 3; d++) o->origin[d] = VO_MIN(A->origin[d], B->origin[d]);
  o->xaxis[0] = VO_MAX(VO_MAX(A->origin[0] + A->xaxis[0], B->origin[0] + B->xaxis[0]), o->origin[0]);
  o->yaxis[1] = VO_MAX(VO_MAX(A->origin[1] + A->yaxis[1], B->origin[1] + B->yaxis[1]), o->origin[1]);
  o->zaxis[2] = VO_MAX(VO_MAX(A->origin[2] + A->zaxis[2], B->origin[2] + B->zaxis[2]), o->origin[2]);
  o->xaxis[0] -= o->origin[0];
  o->yaxis[1] -= o->origin[1];
  o->zaxis[2] -= o->origin[2];
  o->xsize = (int)VO_MAX(dot3d(o->xaxis, A->xaxis) * A->xsize / dot3d(A->xaxis, A->xaxis),
                        dot3d(o->xaxis, B->xaxis) * B->xsize / dot3d(B->xaxis, B->xaxis));
  o->ysize = (int)VO_MAX(dot3d(o->yaxis, A->yaxis) * A->ysize / dot3d(A->yaxis, A->yaxis),
                        dot3d(o->yaxis, B->yaxis) * B->ysize / dot3d(B->yaxis, B->yaxis));
  o->zsize = (int)VO_MAX(dot3d(o->zaxis, A->zaxis) * A->zsize / dot3d(A->zaxis, A->zaxis),
                        dot3d(o->zaxis, B->zaxis) * B->zsize / dot3d(B->zaxis, B->zaxis));
}

/* ------------------------------------------------------------------------------------------
 * Cross correlation [PINNED]  MDFF.C:55-151
 * ------------------------------------------------------------------------------------------ */
int vo_cc(const vo_grid *A, const float *a, const vo_grid *B, const float *b, double thr, double *cc,
          long *nused, double sums[5], int nthreads) {
  vo_grid nv;
  vo_init_from_intersection(A, B, &nv); /* MDFF.C:132 */
  long volsz = (long)(nv.xsize * nv.ysize * nv.zsize); /* MDFF.C:134 (int multiply) */
  float invA[16], invB[16];
  vo_grid_inverse(A, 0, invA);
  vo_grid_inverse(B, 0, invB);

  double sA = 0.0, sB = 0.0, sAA = 0.0, sBB = 0.0, sAB = 0.0;
  long n = 0;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static, 16384) num_threads(nthreads) reduction(+ : sA, sB, sAA, sBB, sAB, n) if (nthreads > 1)
  for (long xi = 0; xi < volsz; xi++) {
    float ix, iy, iz;
    vo_voxel_coord_int(&nv, (int)xi, &ix, &iy, &iz);             /* MDFF.C:69 */
    float vA = interp_from_coord_inv(A, a, invA, ix, iy, iz, 0);  /* MDFF.C:70 */
    float vB = interp_from_coord_inv(B, b, invB, ix, iy, iz, 0);  /* MDFF.C:71 */
    if (!vo_isnan(vB) && !vo_isnan(vA)) {
      if (vA >= thr) { /* MDFF.C:77: float compared against the double threshold */
        sA += vA;
        sB += vB;
        sAA += vA * vA; /* float product, promoted (MDFF.C:80) */
        sBB += vB * vB;
        sAB += vA * vB;
        n++;
      }
    }
  }
  int size = (int)n; /* MDFF.C:143 */
  double aMean = sA / size;
  double bMean = sB / size;
  double stdA = sqrt(sAA / size - aMean * aMean);
  double stdB = sqrt(sBB / size - bMean * bMean);
  *cc = (sAB - size * aMean * bMean) / (size * stdA * stdB); /* MDFF.C:149 */
  if (nused) *nused = n;
  if (sums) { sums[0] = sA; sums[1] = sB; sums[2] = sAA; sums[3] = sBB; sums[4] = sAB; }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * `voltool cc -calcdiffmap / -calcspatialccmap` [RESTATEMENT, UNPINNED]
 * usage TclMDFF.C:180-184, plumbing :555-594.  The maps come out of vmd_cuda_compare_sel_refmap
 * (CUDAMDFF.cu), which is not in the snapshot; only their meaning is stated ("a difference map
 * between the simulated and target densities", "a spatial map of the correlations between 8x8
 * voxel regions").  Definition adopted here and by vt_cc_maps:
 *   G        = init_from_intersection(A, B), the grid cc_threaded walks (MDFF.C:132)
 *   a_i, b_i = the two interpolated samples of MDFF.C:70-71 at voxel i of G
 *   diff[i]  = a_i - b_i on G (NaN where either sample is NaN)
 *   tile t   = 8 x 8 x 8 voxels of G (ragged at the upper faces); its five sums and count run over
 *              the voxels that pass MDFF.C:73-77, added in voxel order (x fastest) inside the tile
 *   spatial[t] = the expression of MDFF.C:143-149 on the tile's sums; 0 for a tile without voxels
 *   spatial grid: ceil(n/8) tiles per axis, origin of G, one cell = 8 cells of G
 *   cc       = MDFF.C:143-149 on the tile sums added in tile order (x fastest): cc_threaded's value
 *              up to the order of its double additions (which its threads do not fix either)
 * ------------------------------------------------------------------------------------------ */
static double cc_expr(const double s[5], long n) { /* MDFF.C:143-149 */
  int size = (int)n;
  double aMean = s[0] / size, bMean = s[1] / size;
  double stdA = sqrt(s[2] / size - aMean * aMean);
  double stdB = sqrt(s[3] / size - bMean * bMean);
  return (s[4] - size * aMean * bMean) / (size * stdA * stdB);
}

void vo_spatial_grid(const vo_grid *G, vo_grid *sg) {
  float xd[3], yd[3], zd[3];
  vo_cell_axes(G, xd, yd, zd);
  *sg = *G;
  sg->xsize = (G->xsize + 7) / 8; sg->ysize = (G->ysize + 7) / 8; sg->zsize = (G->zsize + 7) / 8;
  for (int k = 0; k < 3; k++) {
    sg->xaxis[k] = 8.0 * (double)xd[k] * (double)(sg->xsize > 1 ? sg->xsize - 1 : 1);
    sg->yaxis[k] = 8.0 * (double)yd[k] * (double)(sg->ysize > 1 ? sg->ysize - 1 : 1);
    sg->zaxis[k] = 8.0 * (double)zd[k] * (double)(sg->zsize > 1 ? sg->zsize - 1 : 1);
  }
}

int vo_cc_maps(const vo_grid *A, const float *a, const vo_grid *B, const float *b, double thr, double *cc,
               vo_grid *dg, float **diff, vo_grid *sg, float **spatial) {
  vo_grid nv;
  vo_init_from_intersection(A, B, &nv);
  *dg = nv;
  vo_spatial_grid(&nv, sg);
  long n = vo_gridsize(&nv), nt = vo_gridsize(sg);
  float *d = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  float *sp = (float *)malloc(sizeof(float) * (size_t)(nt > 0 ? nt : 1));
  float invA[16], invB[16];
  vo_grid_inverse(A, 0, invA);
  vo_grid_inverse(B, 0, invB);
  double tot[5] = {0, 0, 0, 0, 0};
  long ntot = 0;
  for (int tz = 0; tz < sg->zsize; tz++)
    for (int ty = 0; ty < sg->ysize; ty++)
      for (int tx = 0; tx < sg->xsize; tx++) {
        double s[5] = {0, 0, 0, 0, 0};
        long cnt = 0;
        for (int lz = 0; lz < 8 && tz * 8 + lz < nv.zsize; lz++)
          for (int ly = 0; ly < 8 && ty * 8 + ly < nv.ysize; ly++)
            for (int lx = 0; lx < 8 && tx * 8 + lx < nv.xsize; lx++) {
              long i = ((long)(tz * 8 + lz) * nv.ysize + (ty * 8 + ly)) * nv.xsize + (tx * 8 + lx);
              float x, y, z;
              vo_voxel_coord(&nv, i, &x, &y, &z);
              float vA = interp_from_coord_inv(A, a, invA, x, y, z, 0);
              float vB = interp_from_coord_inv(B, b, invB, x, y, z, 0);
              d[i] = vA - vB;
              if (!vo_isnan(vB) && !vo_isnan(vA) && vA >= thr) {
                s[0] += vA; s[1] += vB; s[2] += vA * vA; s[3] += vB * vB; s[4] += vA * vB;
                cnt++;
              }
            }
        sp[((long)tz * sg->ysize + ty) * sg->xsize + tx] = cnt > 0 ? (float)cc_expr(s, cnt) : 0.0f;
        for (int k = 0; k < 5; k++) tot[k] += s[k];
        ntot += cnt;
      }
  *cc = cc_expr(tot, ntot);
  *diff = d;
  *spatial = sp;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Two-map ops [PINNED]  Voltool.C:249-377
 * ------------------------------------------------------------------------------------------ */
void vo_binary(int op, int interp, int use_union, const vo_grid *A, const float *a, const vo_grid *B,
               const float *b, vo_grid *og, float **out) {
  if (use_union) vo_init_from_union(A, B, og);
  else vo_init_from_intersection(A, B, og);
  long n = vo_gridsize(og);
  float *o = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  float iA0[16], iB0[16], iA1[16], iB1[16];
  vo_grid_inverse(A, 0, iA0);
  vo_grid_inverse(B, 0, iB0);
  vo_grid_inverse(A, 1, iA1);
  vo_grid_inverse(B, 1, iB1);
  for (long i = 0; i < n; i++) {
    float x, y, z;
    vo_voxel_coord(og, i, &x, &y, &z);
    if (op == 2) { /* multiply: NaN-returning lookups, Voltool.C:303-350 */
      float vA, vB;
      if (interp) {
        vA = interp_from_coord_inv(A, a, iA0, x, y, z, 0);
        vB = interp_from_coord_inv(B, b, iB0, x, y, z, 0);
      } else {
        vA = value_from_coord_inv(A, a, iA1, x, y, z, 0);
        vB = value_from_coord_inv(B, b, iB1, x, y, z, 0);
      }
      if (use_union) {
        if (!vo_isnan(vA) && !vo_isnan(vB)) o[i] = vA * vB;
        else if (!vo_isnan(vA) && vo_isnan(vB)) o[i] = vA;
        else if (vo_isnan(vA) && !vo_isnan(vB)) o[i] = vB;
        else o[i] = 0.f;
      } else {
        o[i] = vA * vB;
      }
    } else {
      float vA, vB;
      if (interp) {
        vA = interp_from_coord_inv(A, a, iA0, x, y, z, 1);
        vB = interp_from_coord_inv(B, b, iB0, x, y, z, 1);
      } else {
        vA = value_from_coord_inv(A, a, iA1, x, y, z, 1);
        vB = value_from_coord_inv(B, b, iB1, x, y, z, 1);
      }
      if (op == 0) o[i] = vA + vB;                     /* Voltool.C:266-271 */
      else if (op == 1) o[i] = vA - vB;                /* Voltool.C:293-298 */
      else o[i] = (float)((vA + vB) * 0.5);            /* Voltool.C:369-374: double 0.5 */
    }
  }
  *out = o;
}

/* ------------------------------------------------------------------------------------------
 * Stats + cache [PINNED modulo the minmax externs, which are not in the snapshot]
 * ------------------------------------------------------------------------------------------ */
void vo_cache_init(vo_cache *c) { memset(c, 0, sizeof(*c)); }

/* utilities.h:108-112 externs [UNPINNED]: sequential min/max, double-precision sum */
void vo_minmax_1fv(const float *f, long n, float *mn, float *mx) {
  if (n < 1) return;
  float lo = f[0], hi = f[0];
  for (long i = 1; i < n; i++) {
    if (f[i] < lo) lo = f[i];
    if (f[i] > hi) hi = f[i];
  }
  *mn = lo; *mx = hi;
}

void vo_minmaxmean_1fv(const float *f, long n, float *mn, float *mx, float *mean) {
  if (n < 1) return;
  float lo = f[0], hi = f[0];
  double s = 0.0;
  for (long i = 0; i < n; i++) {
    if (f[i] < lo) lo = f[i];
    if (f[i] > hi) hi = f[i];
    s += f[i];
  }
  *mn = lo; *mx = hi; *mean = (float)(s / n);
}

/* VolumetricData.C:1047-1074 */
static void compute_minmaxmean(const vo_grid *g, const float *d, vo_cache *c) {
  c->cmin = 0; c->cmax = 0; c->cmean = 0;
  long n = vo_gridsize(g);
  if (n > 0) vo_minmaxmean_1fv(d, n, &c->cmin, &c->cmax, &c->cmean);
  c->mean_valid = 1; c->minmax_valid = 1;
}
static void compute_minmax(const vo_grid *g, const float *d, vo_cache *c) {
  c->cmin = 0; c->cmax = 0;
  long n = vo_gridsize(g);
  if (n > 0) vo_minmax_1fv(d, n, &c->cmin, &c->cmax);
  c->minmax_valid = 1;
}
/* VolumetricData.C:1109-1124 */
static void compute_mean(const vo_grid *g, const float *d, vo_cache *c) {
  double m = 0.0;
  long sz = vo_gridsize(g);
  for (long i = 0; i < sz; i++) m += d[i];
  m /= sz;
  c->cmean = (float)m;
  c->mean_valid = 1;
}
/* VolumetricData.C:1035-1044 */
void vo_datarange(const vo_grid *g, const float *d, vo_cache *c, float *mn, float *mx) {
  if (!c->minmax_valid && !c->mean_valid) compute_minmaxmean(g, d, c);
  else if (!c->minmax_valid) compute_minmax(g, d, c);
  *mn = c->cmin; *mx = c->cmax;
}
/* VolumetricData.C:1078-1086 */
float vo_mean(const vo_grid *g, const float *d, vo_cache *c) {
  if (!c->mean_valid && !c->minmax_valid) compute_minmaxmean(g, d, c);
  else if (!c->mean_valid) compute_mean(g, d, c);
  return c->cmean;
}
/* VolumetricData.C:1095-1100, 1127-1147 */
float vo_sigma(const vo_grid *g, const float *d, vo_cache *c) {
  if (!c->sigma_valid) {
    double s = 0.0;
    long sz = vo_gridsize(g);
    float mymean = vo_mean(g, d, c);
    for (long i = 0; i < sz; i++) {
      float delta = d[i] - mymean;
      s += delta * delta;
    }
    s /= sz;
    s = sqrt(s);
    c->csigma = (float)s;
    c->sigma_valid = 1;
  }
  return c->csigma;
}
static void inval_minmax(vo_cache *c) { c->minmax_valid = 0; c->cmin = 0; c->cmax = 0; }
static void inval_mean(vo_cache *c) { c->mean_valid = 0; c->cmean = 0; }
static void inval_sigma(vo_cache *c) { c->sigma_valid = 0; c->csigma = 0; }

/* Voltool.C:229-247: float accumulators, double mass, strictly sequential */
void vo_vol_com(const vo_grid *g, const float *d, float com[3]) {
  com[0] = com[1] = com[2] = 0.0f;
  double mass = 0.0;
  long n = vo_gridsize(g);
  for (int i = 0; i < n; i++) {
    float m = d[i], ix, iy, iz;
    vo_voxel_coord(g, i, &ix, &iy, &iz);
    mass = mass + m;
    float cx = m * ix, cy = m * iy, cz = m * iz;
    com[0] = com[0] + cx; com[1] = com[1] + cy; com[2] = com[2] + cz;
  }
  float scale = (float)(1.0 / mass);
  com[0] = scale * com[0]; com[1] = scale * com[1]; com[2] = scale * com[2];
}

/* Voltool.C:379-391: the origin goes through float (the map is snapped to float precision) */
void vo_vol_moveto(vo_grid *g, const float com[3], const float pos[3]) {
  float origin[3] = {(float)g->origin[0], (float)g->origin[1], (float)g->origin[2]};
  for (int k = 0; k < 3; k++) {
    float t = pos[k] - com[k];
    origin[k] = origin[k] + t;
  }
  for (int k = 0; k < 3; k++) g->origin[k] = origin[k];
}

/* Voltool.C:399-438: origin = R*origin + t in float; axes = R*axis (double * float products,
   summed in double, stored through float: vectrans, utilities.h:244-248) */
void vo_vol_move(vo_grid *g, const float mat[16]) {
  float origin[3] = {(float)g->origin[0], (float)g->origin[1], (float)g->origin[2]};
  float np[3];
  np[0] = origin[0] * mat[0] + origin[1] * mat[4] + origin[2] * mat[8];
  np[1] = origin[0] * mat[1] + origin[1] * mat[5] + origin[2] * mat[9];
  np[2] = origin[0] * mat[2] + origin[1] * mat[6] + origin[2] * mat[10];
  for (int k = 0; k < 3; k++) { float s = np[k] + mat[12 + k]; g->origin[k] = s; }
  double *axes[3] = {g->xaxis, g->yaxis, g->zaxis};
  for (int a = 0; a < 3; a++) {
    double v[3] = {axes[a][0], axes[a][1], axes[a][2]};
    float r[3];
    r[0] = (float)(v[0] * mat[0] + v[1] * mat[4] + v[2] * mat[8]);
    r[1] = (float)(v[0] * mat[1] + v[1] * mat[5] + v[2] * mat[9]);
    r[2] = (float)(v[0] * mat[2] + v[1] * mat[6] + v[2] * mat[10]);
    for (int k = 0; k < 3; k++) axes[a][k] = r[k];
  }
}

/* Voltool.C:443-461.  The reference indexes bins[nbins] for the maximum voxel (out of bounds);
   here that count is dropped instead of corrupting memory. */
void vo_histogram(const vo_grid *g, const float *d, vo_cache *c, int nbins, long *bins, float *midpts) {
  float mn, mx;
  vo_datarange(g, d, c, &mn, &mx);
  double binwidth = (mx - mn) / nbins;
  double binwidthinv = 1 / binwidth;
  memset(bins, 0, nbins * sizeof(long));
  long n = vo_gridsize(g);
  for (long i = 0; i < n; i++) {
    long k = (long)((d[i] - mn) * binwidthinv);
    if (k >= 0 && k < nbins) bins[k]++;
  }
  for (int j = 0; j < nbins; j++) midpts[j] = (float)(mn + (0.5 * binwidth) + (j * binwidth));
}

/* ------------------------------------------------------------------------------------------
 * Unary streams [PINNED]  VolumetricData.C:634-714, 937-969, 999-1025
 * ------------------------------------------------------------------------------------------ */
void vo_clamp(const vo_grid *g, float *d, vo_cache *c, float lo, float hi) {
  long n = vo_gridsize(g);
  for (long i = 0; i < n; i++) {
    if (d[i] < lo) d[i] = lo;
    else if (d[i] > hi) d[i] = hi;
  }
  if (c->cmin < lo) c->cmin = lo;
  if (c->cmax > hi) c->cmax = hi;
  inval_mean(c); inval_sigma(c);
}

void vo_scale_by(const vo_grid *g, float *d, vo_cache *c, float ff) {
  long n = vo_gridsize(g);
  for (long i = 0; i < n; i++) d[i] *= ff;
  if (ff >= 0.0) { c->cmin *= ff; c->cmax *= ff; }
  else { float t = c->cmin; c->cmin = c->cmax * ff; c->cmax = t * ff; }
  inval_mean(c); inval_sigma(c);
}

void vo_scalar_add(const vo_grid *g, float *d, vo_cache *c, float ff) {
  long n = vo_gridsize(g);
  for (long i = 0; i < n; i++) d[i] += ff;
  c->cmin += ff; c->cmax += ff;
  inval_mean(c); inval_sigma(c);
}

void vo_rescale_range(const vo_grid *g, float *d, vo_cache *c, float lo, float hi) {
  float newvrange = hi - lo;
  float oldvrange = c->cmax - c->cmin;
  float ratio = newvrange / oldvrange;
  long n = vo_gridsize(g);
  for (long i = 0; i < n; i++) d[i] = lo + ratio * (d[i] - c->cmin);
  c->cmin = lo; c->cmax = hi;
  inval_mean(c); inval_sigma(c);
}

void vo_sigma_scale(const vo_grid *g, float *d, vo_cache *c) {
  float oldmean = vo_mean(g, d, c);
  float oldsigma = vo_sigma(g, d, c);
  float inv = 1.0f / oldsigma;
  long n = vo_gridsize(g);
  for (long i = 0; i < n; i++) { d[i] -= oldmean; d[i] *= inv; }
  inval_minmax(c); inval_mean(c); inval_sigma(c);
}

void vo_binmask(const vo_grid *g, float *d, vo_cache *c, float thr) {
  long n = vo_gridsize(g);
  for (long i = 0; i < n; i++) { float t = d[i]; d[i] = (t > thr) ? 1 : 0; }
  inval_minmax(c); inval_mean(c); inval_sigma(c);
}

void vo_mdff_potential(const vo_grid *g, float *d, vo_cache *c, double threshold) {
  float threshinvert = (float)(threshold * -1);
  float mininvert = (c->cmax * -1);
  float oldvrange = threshinvert - mininvert;
  float ratio = 1 / oldvrange;
  long n = vo_gridsize(g);
  for (long i = 0; i < n; i++) {
    if (d[i] < threshold) d[i] = (float)threshold;
    d[i] = d[i] * -1;
    d[i] = ratio * (d[i] - mininvert);
  }
  c->cmin = 0; c->cmax = 1;
  inval_mean(c); inval_sigma(c);
}

/* ------------------------------------------------------------------------------------------
 * Resizing ops [PINNED]
 * ------------------------------------------------------------------------------------------ */
/* VolumetricData.C:542-609 */
void vo_pad(vo_grid *g, const float *d, float **out, int pxm, int pxp, int pym, int pyp, int pzm, int pzp) {
  int xs = g->xsize, ys = g->ysize, zs = g->zsize;
  double xd[3], yd[3], zd[3];
  for (int k = 0; k < 3; k++) {
    xd[k] = g->xaxis[k] / (xs - 1);
    yd[k] = g->yaxis[k] / (ys - 1);
    zd[k] = g->zaxis[k] / (zs - 1);
  }
  int nx = VO_MAX(1, xs + pxm + pxp), ny = VO_MAX(1, ys + pym + pyp), nz = VO_MAX(1, zs + pzm + pzp);
  long ng = (long)nx * ny * nz;
  float *o = (float *)calloc((size_t)ng, sizeof(float));
  int sx = VO_MAX(0, pxm), sy = VO_MAX(0, pym), sz = VO_MAX(0, pzm);
  int ex = VO_MIN(nx, xs + pxm), ey = VO_MIN(ny, ys + pym), ez = VO_MIN(nz, zs + pzm);
  for (int gz = sz; gz < ez; gz++)
    for (int gy = sy; gy < ey; gy++) {
      long oldaddr = (gy - pym) * (long)xs + (gz - pzm) * (long)xs * (long)ys - pxm;
      long newaddr = gy * (long)nx + gz * (long)nx * (long)ny;
      for (int gx = sx; gx < ex; gx++) o[gx + newaddr] = d[gx + oldaddr];
    }
  g->xsize = nx; g->ysize = ny; g->zsize = nz;
  /* vec_scaled_add(double*, float, const double*): utilities.h:228-232 */
  float fx = (float)(pxm + pxp), fy = (float)(pym + pyp), fz = (float)(pzm + pzp);
  for (int k = 0; k < 3; k++) g->xaxis[k] += fx * xd[k];
  for (int k = 0; k < 3; k++) g->yaxis[k] += fy * yd[k];
  for (int k = 0; k < 3; k++) g->zaxis[k] += fz * zd[k];
  float mx = (float)(-pxm), my = (float)(-pym), mz = (float)(-pzm);
  for (int k = 0; k < 3; k++) g->origin[k] += mx * xd[k];
  for (int k = 0; k < 3; k++) g->origin[k] += my * yd[k];
  for (int k = 0; k < 3; k++) g->origin[k] += mz * zd[k];
  *out = o;
}

/* VolumetricData.C:615-631 */
void vo_crop(vo_grid *g, const float *d, float **out, double minx, double miny, double minz, double maxx,
             double maxy, double maxz) {
  double xd0 = g->xaxis[0] / (g->xsize - 1);
  double yd1 = g->yaxis[1] / (g->ysize - 1);
  double zd2 = g->zaxis[2] / (g->zsize - 1);
  int pxm = (int)((g->origin[0] - minx) / xd0);
  int pym = (int)((g->origin[1] - miny) / yd1);
  int pzm = (int)((g->origin[2] - minz) / zd2);
  int pxp = (int)((maxx - g->origin[0] - g->xaxis[0]) / xd0);
  int pyp = (int)((maxy - g->origin[1] - g->yaxis[1]) / yd1);
  int pzp = (int)((maxz - g->origin[2] - g->zaxis[2]) / zd2);
  vo_pad(g, d, out, pxm, pxp, pym, pyp, pzm, pzp);
}

/* VolumetricData.C:720-769 */
void vo_downsample(vo_grid *g, const float *d, float **out) {
  int xs = g->xsize, ys = g->ysize, zs = g->zsize;
  int nx = xs / 2, ny = ys / 2, nz = zs / 2;
  long ng = (long)nx * ny * nz;
  float *o = (float *)malloc(sizeof(float) * (size_t)(ng > 0 ? ng : 1));
  long shift[8] = {0, 1, xs, xs + 1, (long)xs * ys, (long)xs * ys + 1, (long)xs * ys + xs, (long)xs * ys + xs + 1};
  const double eighth = 1.0 / 8.0;
  for (int gz = 0; gz < nz; gz++)
    for (int gy = 0; gy < ny; gy++) {
      long oldaddr = gy * (long)xs + gz * (long)xs * (long)ys;
      long newaddr = gy * (long)nx + gz * (long)nx * (long)ny;
      for (int gx = 0; gx < nx; gx++) {
        long n = 2 * (gx + oldaddr);
        double Z = 0.0;
        for (int j = 0; j < 8; j++) Z += d[n + shift[j]];
        o[gx + newaddr] = (float)(Z * eighth);
      }
    }
  g->xsize = nx; g->ysize = ny; g->zsize = nz;
  /* VolumetricData.C:750-757: scale computed from the NEW sizes with integer division */
  double xscale = 0.5 * (g->xsize) / (g->xsize / 2);
  double yscale = 0.5 * (g->ysize) / (g->ysize / 2);
  double zscale = 0.5 * (g->zsize) / (g->zsize / 2);
  for (int i = 0; i < 3; i++) { g->xaxis[i] *= xscale; g->yaxis[i] *= yscale; g->zaxis[i] *= zscale; }
  *out = o;
}

/* VolumetricData.h:224-232 */
static float cubic_interp(float y0, float y1, float y2, float y3, float mu) {
  float mu2 = mu * mu;
  float a0 = y3 - y2 - y0 + y1;
  float a1 = y0 - y1 - a0;
  float a2 = y2 - y0;
  float a3 = y1;
  return (a0 * mu * mu2 + a1 * mu2 + a2 * mu + a3);
}

/* VolumetricData.C:773-932 (requires every size >= 3, as the reference implicitly does) */
void vo_supersample(vo_grid *g, const float *d, float **out) {
  int xs = g->xsize, ys = g->ysize, zs = g->zsize;
  int nx = xs * 2 - 1, ny = ys * 2 - 1, nz = zs * 2 - 1;
  long xy = (long)xs * ys, nxy = (long)nx * ny;
  float *o = (float *)malloc(sizeof(float) * (size_t)((long)nx * ny * nz));
  int gx, gy, gz;
  for (gz = 0; gz < zs; gz++)
    for (gy = 0; gy < ys; gy++) {
      long op = (long)gy * xs + (long)gz * xy;
      long np = 2L * gy * nx + 2L * gz * nxy;
      for (gx = 0; gx < xs; gx++) o[2 * gx + np] = d[gx + op];
    }
  /* x direction */
  for (gz = 0; gz < zs; gz++)
    for (gy = 0; gy < ys; gy++) {
      long np = 2L * gy * nx + 2L * gz * nxy;
      for (gx = 1; gx < xs - 2; gx++)
        o[2 * gx + 1 + np] = cubic_interp(o[(2 * gx - 2) + np], o[(2 * gx) + np], o[(2 * gx + 2) + np],
                                          o[(2 * gx + 4) + np], 0.5f);
    }
  for (gz = 0; gz < zs; gz++)
    for (gy = 0; gy < ys; gy++) {
      long np = 2L * gy * nx + 2L * gz * nxy;
      o[1 + np] = cubic_interp(o[0 + np], o[0 + np], o[2 + np], o[4 + np], 0.5f);
      o[2 * (xs - 2) + 1 + np] = cubic_interp(o[2 * (xs - 2) - 2 + np], o[2 * (xs - 2) + np],
                                              o[2 * (xs - 2) + 2 + np], o[2 * (xs - 2) + 2 + np], 0.5f);
    }
  /* y direction */
  for (gz = 0; gz < zs; gz++)
    for (gy = 1; gy < ys - 2; gy++) {
      long nzp = 2L * gz * nxy;
      for (gx = 0; gx < nx; gx++)
        o[gx + (2 * gy + 1) * (long)nx + nzp] =
            cubic_interp(o[gx + (2 * gy - 2) * (long)nx + nzp], o[gx + (2 * gy) * (long)nx + nzp],
                         o[gx + (2 * gy + 2) * (long)nx + nzp], o[gx + (2 * gy + 4) * (long)nx + nzp], 0.5f);
    }
  for (gz = 0; gz < zs; gz++) {
    long nzp = 2L * gz * nxy;
    for (gx = 0; gx < nx; gx++) {
      o[gx + 1 * (long)nx + nzp] =
          cubic_interp(o[gx + nzp], o[gx + nzp], o[gx + 2 * (long)nx + nzp], o[gx + 4 * (long)nx + nzp], 0.5f);
      o[gx + (2 * (ys - 2) + 1) * (long)nx + nzp] =
          cubic_interp(o[gx + (2 * (ys - 2) - 2) * (long)nx + nzp], o[gx + 2 * (ys - 2) * (long)nx + nzp],
                       o[gx + (2 * (ys - 2) + 2) * (long)nx + nzp], o[gx + (2 * (ys - 2) + 2) * (long)nx + nzp],
                       0.5f);
    }
  }
  /* z direction */
  for (gy = 0; gy < ny; gy++)
    for (gz = 1; gz < zs - 2; gz++) {
      long np = gy * (long)nx + (2L * gz + 1L) * nxy;
      for (gx = 0; gx < nx; gx++) {
        long r = gx + gy * (long)nx;
        o[gx + np] = cubic_interp(o[r + (2 * gz - 2) * nxy], o[r + (2 * gz) * nxy], o[r + (2 * gz + 2) * nxy],
                                  o[r + (2 * gz + 4) * nxy], 0.5f);
      }
    }
  for (gy = 0; gy < ny; gy++)
    for (gx = 0; gx < nx; gx++) {
      long r = gx + gy * (long)nx;
      o[r + 1 * nxy] = cubic_interp(o[r], o[r], o[r + 2 * nxy], o[r + 4 * nxy], 0.5f);
      o[r + (2 * (zs - 2) + 1) * nxy] =
          cubic_interp(o[r + (2 * (zs - 2) - 2) * nxy], o[r + 2 * (zs - 2) * nxy], o[r + (2 * (zs - 2) + 2) * nxy],
                       o[r + (2 * (zs - 2) + 2) * nxy], 0.5f);
    }
  g->xsize = nx; g->ysize = ny; g->zsize = nz;
  *out = o;
}

/* ------------------------------------------------------------------------------------------
 * Gaussian blur [UNPINNED: GaussianBlur.C / Segmentation.h are not in the snapshot].
 * Call site VolumetricData.C:977-997; sigma is in voxels (usage text TclVoltool.C:1551).
 * Definition adopted here: separable x -> y -> z; taps k = -R..R with
 * R = ceil(3 sigma) (>= 1); weights expf(-k^2 / (2 sigma^2)) normalised by their float sum;
 * clamp-to-edge borders; each pass accumulates acc = fmaf(w[k], in[i+k], acc) in tap order and
 * rounds to f32 between passes.
 * ------------------------------------------------------------------------------------------ */
int vo_blur_kernel(double sigma, float *w, int maxw) {
  int R = (int)ceil(3.0 * sigma);
  if (R < 1) R = 1;
  if (2 * R + 1 > maxw) return -1;
  float sf = (float)sigma;
  float denom = 2.0f * sf * sf;
  float sum = 0.0f;
  for (int k = -R; k <= R; k++) {
    float v = expf(-(float)(k * k) / denom);
    w[k + R] = v;
    sum += v;
  }
  for (int k = 0; k <= 2 * R; k++) w[k] = w[k] / sum;
  return R;
}

static void blur_axis(const float *in, float *out, int nx, int ny, int nz, int axis, const float *w, int R) {
  long sx = 1, sy = nx, sz = (long)nx * ny;
  long stride = axis == 0 ? sx : (axis == 1 ? sy : sz);
  int n = axis == 0 ? nx : (axis == 1 ? ny : nz);
#pragma omp parallel for schedule(static)
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        long base = x * sx + y * sy + z * sz;
        int i = axis == 0 ? x : (axis == 1 ? y : z);
        float acc = 0.0f;
        for (int k = -R; k <= R; k++) {
          int j = i + k;
          if (j < 0) j = 0;
          if (j > n - 1) j = n - 1;
          acc = fmaf(w[k + R], in[base + (long)(j - i) * stride], acc);
        }
        out[base] = acc;
      }
}

void vo_gaussian_blur(const vo_grid *g, const float *d, float **out, double sigma) {
  long n = vo_gridsize(g);
  float *a = (float *)malloc(sizeof(float) * (size_t)n);
  float *b = (float *)malloc(sizeof(float) * (size_t)n);
  float w[1024];
  int R = (sigma > 0.0) ? vo_blur_kernel(sigma, w, 1024) : -1;
  if (R < 0) { memcpy(a, d, sizeof(float) * (size_t)n); free(b); *out = a; return; }
  blur_axis(d, a, g->xsize, g->ysize, g->zsize, 0, w, R);
  blur_axis(a, b, g->xsize, g->ysize, g->zsize, 1, w, R);
  blur_axis(b, a, g->xsize, g->ysize, g->zsize, 2, w, R);
  free(b);
  *out = a;
}

/* ------------------------------------------------------------------------------------------
 * QuickSurf density map [UNPINNED: QuickSurf.C is not in the snapshot].
 * Call sites TclMDFF.C:154-157, 496-499; TclVoltool.C:119-122.  Definition adopted here
 * (SURVEY 8c):
 *   grid:  bbox of the atoms, padded by max(p, 0.65*sqrt(4/3*pi*p^3)) with p = radscale*maxrad*1.70;
 *          numvoxels = ceil(extent / gspacing); axes = (numvoxels-1)*gspacing; origin = padded min.
 *   splat: per atom sigma = r*radscale, cutoff = gausslim*sigma; index box
 *          [(int)(p/gs - cutoff/gs), (int)(p/gs + cutoff/gs)] clamped to the grid; rows with
 *          dy^2+dz^2 >= cutoff^2 skipped (x bounded by the box only); each voxel accumulates
 *          exp2f(r^2 * arinv), arinv = -(1/(2 sigma^2))*log2(e), in f32, in atom order.
 *   Voxel offsets are computed directly (d = idx*gs - p), not by a running sum, so the value
 *   does not depend on the traversal; exp2f is libm's (the upstream CPU code uses a
 *   polynomial approximation whose constants could not be recovered).
 *   maxrad is taken over the atoms passed in.
 * ------------------------------------------------------------------------------------------ */
float vo_gausslim_for_quality(int quality) {
  switch (quality) {
    case 3: return 4.0f;
    case 2: return 3.0f;
    case 1: return 2.5f;
    default: return 2.0f;
  }
}

/* TclMDFF.C:125-135, 397-401, 492-494; TclVoltool.C:81-88 */
void vo_sim_policy(double resolution, double gspacing_in, float *radscale, double *gspacing, int *quality) {
  *radscale = (float)(.2 * resolution);
  *gspacing = gspacing_in;
  if (*gspacing == 0) *gspacing = 1.5 * (*radscale);
  *quality = (resolution >= 9) ? 0 : 3;
}

void vo_sim_grid(const float *pos, const float *radii, long natoms, float radscale, float gspacing, vo_grid *g,
                 float origin_f[3]) {
  float mn[3] = {pos[0], pos[1], pos[2]}, mx[3] = {pos[0], pos[1], pos[2]};
  float maxrad = radii[0];
  for (long i = 0; i < natoms; i++) {
    for (int k = 0; k < 3; k++) {
      float v = pos[3 * i + k];
      if (v < mn[k]) mn[k] = v;
      if (v > mx[k]) mx[k] = v;
    }
    if (radii[i] > maxrad) maxrad = radii[i];
  }
  float pad = radscale * maxrad * 1.70f;
  float padrad = pad;
  padrad = 0.65f * sqrtf(4.0f / 3.0f * ((float)VO_PI) * padrad * padrad * padrad);
  pad = VO_MAX(pad, padrad);
  int nv[3];
  float ax[3];
  for (int k = 0; k < 3; k++) {
    mn[k] -= pad;
    mx[k] += pad;
    ax[k] = mx[k] - mn[k];
    nv[k] = (int)ceil(ax[k] / gspacing);
    ax[k] = (nv[k] - 1) * gspacing;
    origin_f[k] = mn[k];
  }
  memset(g, 0, sizeof(*g));
  for (int k = 0; k < 3; k++) g->origin[k] = mn[k];
  g->xaxis[0] = ax[0]; g->yaxis[1] = ax[1]; g->zaxis[2] = ax[2];
  g->xsize = nv[0]; g->ysize = nv[1]; g->zsize = nv[2];
}

#define VO_MLOG2EF 1.44269504088896f

static void splat_range(const float *xyzr, long natoms, float *map, const int nv[3], float radscale, float gs,
                        float gausslim, int zlo, int zhi) {
  const float invgs = 1.0f / gs;
  for (long i = 0; i < natoms; i++) {
    const float *a = xyzr + 4 * i;
    float scaledrad = a[3] * radscale;
    float arinv = -(1.0f / (2.0f * scaledrad * scaledrad)) * VO_MLOG2EF;
    float radlim = gausslim * scaledrad;
    float radlim2 = radlim * radlim;
    radlim *= invgs;
    float tmp;
    tmp = a[0] * invgs;
    int xmin = VO_MAX((int)(tmp - radlim), 0), xmax = VO_MIN((int)(tmp + radlim), nv[0] - 1);
    tmp = a[1] * invgs;
    int ymin = VO_MAX((int)(tmp - radlim), 0), ymax = VO_MIN((int)(tmp + radlim), nv[1] - 1);
    tmp = a[2] * invgs;
    int zmin = VO_MAX((int)(tmp - radlim), 0), zmax = VO_MIN((int)(tmp + radlim), nv[2] - 1);
    if (zmin < zlo) zmin = zlo;
    if (zmax > zhi - 1) zmax = zhi - 1;
    for (int z = zmin; z <= zmax; z++) {
      float dz = z * gs - a[2];
      for (int y = ymin; y <= ymax; y++) {
        float dy = y * gs - a[1];
        float dy2dz2 = dy * dy + dz * dz;
        if (dy2dz2 >= radlim2) continue;
        long addr = (long)z * nv[0] * nv[1] + (long)y * nv[0];
        for (int x = xmin; x <= xmax; x++) {
          float dx = x * gs - a[0];
          float r2 = dx * dx + dy2dz2;
          float mb = r2 * arinv;
          map[addr + x] += exp2f(mb);
        }
      }
    }
  }
}

void vo_sim(const float *pos, const float *radii, long natoms, int quality, float radscale, float gspacing,
            vo_grid *g, float **out, int nthreads) {
  float of[3];
  vo_sim_grid(pos, radii, natoms, radscale, gspacing, g, of);
  int nv[3] = {g->xsize, g->ysize, g->zsize};
  long n = vo_gridsize(g);
  float *map = (float *)calloc((size_t)(n > 0 ? n : 1), sizeof(float));
  float *xyzr = (float *)malloc(sizeof(float) * 4 * (size_t)natoms);
  for (long i = 0; i < natoms; i++) {
    xyzr[4 * i + 0] = pos[3 * i + 0] - of[0];
    xyzr[4 * i + 1] = pos[3 * i + 1] - of[1];
    xyzr[4 * i + 2] = pos[3 * i + 2] - of[2];
    xyzr[4 * i + 3] = radii[i];
  }
  float gausslim = vo_gausslim_for_quality(quality);
  if (nthreads <= 1) {
    splat_range(xyzr, natoms, map, nv, radscale, gspacing, gausslim, 0, nv[2]);
  } else {
    /* z-slabs: per-voxel accumulation order (atom order) is unchanged, so the result is
       bit-identical to the single-thread run */
    int nslab = nthreads * 4;
    if (nslab > nv[2]) nslab = nv[2];
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int s = 0; s < nslab; s++) {
      int zlo = (int)((long)nv[2] * s / nslab), zhi = (int)((long)nv[2] * (s + 1) / nslab);
      splat_range(xyzr, natoms, map, nv, radscale, gspacing, gausslim, zlo, zhi);
    }
  }
  free(xyzr);
  *out = map;
}

/* ------------------------------------------------------------------------------------------
 * Rigid transforms and the fit driver
 * ------------------------------------------------------------------------------------------ */
/* TclVoltool.C:52-59 (selection = all atoms passed) */
void vo_moveby(float *pos, long natoms, const float v[3]) {
  for (long i = 0; i < natoms; i++) {
    pos[3 * i + 0] = pos[3 * i + 0] + v[0];
    pos[3 * i + 1] = pos[3 * i + 1] + v[1];
    pos[3 * i + 2] = pos[3 * i + 2] + v[2];
  }
}

/* measure_move (Measure.C, not in the snapshot) [UNPINNED]: multpoint3d on every selected atom */
void vo_apply_mat(float *pos, long natoms, const float m[16]) {
  for (long i = 0; i < natoms; i++) vo_mat4_multpoint3d(m, pos + 3 * i, pos + 3 * i);
}

/* measure_center (Measure.C, not in the snapshot) [UNPINNED]: weighted mean, float products
   accumulated in double, call sites TclVoltool.C:789, 803 */
void vo_measure_center(const float *pos, const float *w, long natoms, float com[3]) {
  double x = 0, y = 0, z = 0, ws = 0;
  for (long i = 0; i < natoms; i++) {
    float tw = w[i];
    ws += (double)tw;
    x += (double)(tw * pos[3 * i + 0]);
    y += (double)(tw * pos[3 * i + 1]);
    z += (double)(tw * pos[3 * i + 2]);
  }
  com[0] = (float)(x / ws); com[1] = (float)(y / ws); com[2] = (float)(z / ws);
}

/* one candidate of rotate(): TclVoltool.C:157-176, do_rotate :131-136 */
void vo_pose_transform(float *pos, long natoms, const float com[3], int stride, int x, int y, int z) {
  float move1[3] = {-1.0f * com[0], -1.0f * com[1], -1.0f * com[2]};
  float m[16];
  vo_moveby(pos, natoms, move1);
  int amt[3] = {x, y, z};
  for (int ax = 0; ax < 3; ax++) {
    double amount = (stride * amt[ax] * VO_PI / 180.0); /* DEGTORAD, utilities.h:49 */
    vo_mat4_rot(m, ax, (float)amount);
    vo_apply_mat(pos, natoms, m);
  }
  vo_moveby(pos, natoms, com);
}

/* one candidate of an explicit pose list (the translation-search extension of the fit path): the body of
   rotate() (TclVoltool.C:157-176) with the integer `stride * amt` of do_rotate (:131-136) replaced by a
   float number of degrees per axis, then moveby(shift).  shift = {0,0,0} and rot = stride * {x,y,z}
   reproduces vo_pose_transform bit for bit (tests/test_oracle_vs_ref.py). */
void vo_pose_general(float *pos, long natoms, const float com[3], const float rot_deg[3], const float shift[3]) {
  float move1[3] = {-1.0f * com[0], -1.0f * com[1], -1.0f * com[2]};
  float m[16];
  vo_moveby(pos, natoms, move1);
  for (int ax = 0; ax < 3; ax++) {
    double amount = ((double)rot_deg[ax] * VO_PI / 180.0); /* DEGTORAD, utilities.h:49 */
    vo_mat4_rot(m, ax, (float)amount);
    vo_apply_mat(pos, natoms, m);
  }
  vo_moveby(pos, natoms, com);
  vo_moveby(pos, natoms, shift);
}

static vo_cc_fn g_cc_fn = 0;
static void *g_cc_ctx = 0;
void vo_set_cc_backend(vo_cc_fn fn, void *ctx) { g_cc_fn = fn; g_cc_ctx = ctx; }

/* TclVoltool.C:73-129 (CPU branch) */
double vo_calc_cc(const float *pos, const float *radii, long natoms, const vo_grid *T, const float *t,
                  float resolution, int nthreads) {
  float radscale;
  double gspacing = 0;
  double thresholddensity = 0.1;
  float return_cc = 0;
  radscale = (float)(.2 * resolution);
  gspacing = 1.5 * radscale;
  int quality = (resolution >= 9) ? 0 : 3;
  vo_grid sg;
  float *sim = 0;
  vo_sim(pos, radii, natoms, quality, (float)radscale, (float)gspacing, &sg, &sim, nthreads);
  double cc = 0.0;
  if (g_cc_fn) cc = g_cc_fn(&sg, sim, T, t, thresholddensity, g_cc_ctx);
  else vo_cc(&sg, sim, T, t, thresholddensity, &cc, 0, 0, nthreads);
  return_cc += cc;
  free(sim);
  return return_cc;
}

void vo_rotate_table(const float *origin_pos, const float *radii, long natoms, const float com[3], int stride,
                     int max_rot, const vo_grid *T, const float *t, float resolution, float *cc_table,
                     int pose_stride, int pose_first, int nthreads) {
  int total = max_rot * max_rot * max_rot;
  if (pose_stride < 1) pose_stride = 1;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) if (nthreads > 1)
  for (int idx = pose_first; idx < total; idx += pose_stride) {
    float *p = (float *)malloc(sizeof(float) * 3 * (size_t)natoms);
    memcpy(p, origin_pos, sizeof(float) * 3 * (size_t)natoms);
    int x = idx / (max_rot * max_rot), y = (idx / max_rot) % max_rot, z = idx % max_rot;
    vo_pose_transform(p, natoms, com, stride, x, y, z);
    cc_table[idx] = (float)vo_calc_cc(p, radii, natoms, T, t, resolution, 1);
    free(p);
  }
}

/* TclVoltool.C:160-199 driven from a table of candidate CCs */
int vo_rotate_replay(const float *cc_table, int max_rot, float *returncc, int *best_idx) {
  int visited = 0;
  *best_idx = -1;
  for (int x = 0; x < max_rot; x++)
    for (int y = 0; y < max_rot; y++)
      for (int z = 0; z < max_rot; z++) {
        int idx = (x * max_rot + y) * max_rot + z;
        float cc = cc_table[idx];
        visited++;
        if (cc > *returncc) {
          *returncc = cc;
          *best_idx = idx;
        } else if (cc < *returncc) {
          z++;
        }
      }
  return visited;
}

/* ------------------------------------------------------------------------------------------
 * Map-rotating search [RESTATEMENT of dead code]  density_rotate TclVoltool.C:202-279, do_density_rotate
 * :138-149, reset_density :281-311 (caller commented out, :820-891).  Per candidate (x, y, z): move the
 * synthetic map's frame by -com (vol_moveto with pos = 0), rotate origin and axes about X, Y, Z
 * (vol_move, Voltool.C:399-438), recompute its vol_com and move that back onto com, score
 * cc_threaded(target, synth, 0.1) -- the TARGET is the thresholded map here (:252) -- and keep the
 * best by `cc > *returncc` (no skip rule in this loop), then put the frame back.
 * One deliberate deviation: the dead caller passes synthvol's OWN origin/axis arrays as the copies to
 * restore from (:874-877), so its restore (:262-273) is a self-assignment and the rotations would pile
 * up; this restatement restores from true copies taken on entry, which is what the loop is written to do.
 * ------------------------------------------------------------------------------------------ */
static void density_pose(vo_grid *S, const float *s, const float com[3], int stride, const int amt[3]) {
  const float zero[3] = {0, 0, 0};
  float m[16], currcom[3];
  vo_vol_moveto(S, com, zero);                           /* :210 */
  for (int ax = 0; ax < 3; ax++) {                       /* :212-222 */
    double amount = (stride * amt[ax] * VO_PI / 180.0);  /* DEGTORAD, :141 */
    vo_mat4_rot(m, ax, (float)amount);
    vo_vol_move(S, m);
  }
  vo_vol_com(S, s, currcom);                             /* :225 */
  vo_vol_moveto(S, currcom, com);                        /* :226 */
}

int vo_density_rotate(int stride, int max_rot, const float com[3], float *returncc, int bestrot[3], const vo_grid *A,
                      const float *a, vo_grid *S, const float *s, float *cc_table, int nthreads) {
  const vo_grid saved = *S;
  int n = 0;
  for (int x = 0; x < max_rot; x++)
    for (int y = 0; y < max_rot; y++)
      for (int z = 0; z < max_rot; z++) {
        const int amt[3] = {x, y, z};
        density_pose(S, s, com, stride, amt);
        double ccd = 0.0;
        vo_cc(A, a, S, s, 0.1, &ccd, 0, 0, nthreads);    /* density_calc_cc :66-71, :252 */
        float cc = (float)ccd;
        if (cc_table) cc_table[n] = cc;
        n++;
        if (cc > *returncc) { *returncc = cc; bestrot[0] = x; bestrot[1] = y; bestrot[2] = z; }   /* :254-261 */
        *S = saved;                                      /* :262-273 */
      }
  return n;
}

/* the map half of reset_density (:283-300); the atom half (:302-311) is vo_pose_transform */
void vo_density_reset(int stride, const int bestrot[3], vo_grid *S, const float *s, const float synthcom[3],
                      const float com[3]) {
  float move1[3] = {-1.0f * synthcom[0], -1.0f * synthcom[1], -1.0f * synthcom[2]};
  float m[16], currcom[3];
  vo_vol_moveto(S, synthcom, move1);                     /* :285 */
  for (int ax = 0; ax < 3; ax++) {
    double amount = (stride * bestrot[ax] * VO_PI / 180.0);
    vo_mat4_rot(m, ax, (float)amount);
    vo_vol_move(S, m);
  }
  vo_vol_com(S, s, currcom);                             /* :298 */
  vo_vol_moveto(S, currcom, com);                        /* :299 */
}

/* literal rotate(): TclVoltool.C:151-200 */
static int rotate_direct(int stride, int max_rot, const float *com, float *returncc, float *bestpos, const float *radii,
                         long natoms, const vo_grid *T, const float *t, float resolution, const float *origin,
                         float *framepos, int *best_idx, int nthreads) {
  int visited = 0;
  *best_idx = -1;
  for (int x = 0; x < max_rot; x++)
    for (int y = 0; y < max_rot; y++)
      for (int z = 0; z < max_rot; z++) {
        int idx = (x * max_rot + y) * max_rot + z;
        vo_pose_transform(framepos, natoms, com, stride, x, y, z);
        float cc = (float)vo_calc_cc(framepos, radii, natoms, T, t, resolution, nthreads);
        visited++;
        if (cc > *returncc) {
          *returncc = cc;
          *best_idx = idx;
          memcpy(bestpos, framepos, sizeof(float) * 3 * (size_t)natoms);
        } else if (cc < *returncc) {
          z++;
        }
        memcpy(framepos, origin, sizeof(float) * 3 * (size_t)natoms);
      }
  return visited;
}

/* fit(): TclVoltool.C:784-916 */
void vo_fit(float *pos, const float *radii, const float *mass, long natoms, const vo_grid *T, const float *t,
            float resolution, vo_fit_result *res, int direct, int nthreads) {
  size_t nb = sizeof(float) * 3 * (size_t)natoms;
  float com[3], dencom[3] = {0, 0, 0}, move1[3];
  vo_measure_center(pos, mass, natoms, com);           /* :789 */
  vo_vol_com(T, t, dencom);                             /* :799 */
  for (int k = 0; k < 3; k++) move1[k] = dencom[k] - com[k];
  vo_moveby(pos, natoms, move1);                        /* :801 */
  vo_measure_center(pos, mass, natoms, com);           /* :803 */
  float cc = -1;
  float *bestpos = (float *)malloc(nb), *origin = (float *)malloc(nb);
  /* table mode only: what framepos holds when a rotate() call starts, and the score of its
     zero-rotation candidate.  reset_origin() rewrites origin but not framepos (:314-318), so the
     first candidate after a reset re-scores the OLD origin; the literal loop (direct) does this by
     construction, the table replay has to be told. */
  float *frame = (float *)malloc(nb);
  float frame_cc0 = 0.0f;
  memcpy(origin, pos, nb);
  memcpy(bestpos, pos, nb);
  memcpy(frame, pos, nb);
  res->n_evaluated = 0;
  for (int k = 0; k < 3; k++) { res->com[k] = com[k]; res->dencom[k] = dencom[k]; }

  /* :895-909 */
  const int strides[5] = {24, 4, -4, 1, -1};
  const int maxrots[5] = {360 / 24, 24 / 4, 24 / 4, 4 / 1, 4 / 1};
  for (int s = 0; s < 5; s++) {
    int stride = strides[s], m = maxrots[s], bi = -1;
    if (direct) {
      res->n_evaluated += rotate_direct(stride, m, com, &cc, bestpos, radii, natoms, T, t, resolution, origin, pos,
                                        &bi, nthreads);
    } else {
      float *tab = (float *)malloc(sizeof(float) * (size_t)(m * m * m));
      vo_rotate_table(origin, radii, natoms, com, stride, m, T, t, resolution, tab, 1, 0, nthreads);
      float fresh_cc0 = tab[0];
      int stale = memcmp(frame, origin, nb) != 0;
      if (stale) tab[0] = frame_cc0;
      res->n_evaluated += vo_rotate_replay(tab, m, &cc, &bi);
      if (bi >= 0) {
        memcpy(bestpos, (bi == 0 && stale) ? frame : origin, nb);
        vo_pose_transform(bestpos, natoms, com, stride, bi / (m * m), (bi / m) % m, bi % m);
      }
      memcpy(frame, origin, nb);
      frame_cc0 = fresh_cc0;
      free(tab);
    }
    res->stage_best[s] = bi;
    if (s == 0 || s == 2) memcpy(origin, bestpos, nb); /* reset_origin :898, :904 */
  }
  free(frame);
  memcpy(pos, bestpos, nb); /* :911-913 */
  res->best_cc = cc;
  free(bestpos);
  free(origin);
}
