// vt_geom.h -- geometry arithmetic shared VERBATIM by host code (g++ -ffp-contract=off) and device
// code (nvcc --fmad=false): one source, so the per-pose setup kernel of the fit path and the
// host-driven single-map path produce the same bits.  Divisions and square roots are IEEE on
// both sides (nvcc defaults -prec-div=true -prec-sqrt=true).
#pragma once
#include "../../include/voltool_gpu.h"

#if defined(__CUDACC__)
#define VT_HD __host__ __device__ __forceinline__
#else
#define VT_HD inline
#endif

namespace vt {
namespace geom {

VT_HD float absf(float v) { return v < 0.0f ? -v : v; }
VT_HD double maxd(double a, double b) { return a > b ? a : b; }   // MAX/MIN macros of utilities.h
VT_HD double mind(double a, double b) { return a < b ? a : b; }

// Matrix4::inverse restatement (see vt_host.h): diagonal-pivot Gauss-Jordan on [M | I] in float
VT_HD bool invert4(float m[16]) {
  float aug[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      aug[r][c] = m[4 * r + c];
      aug[r][4 + c] = r == c ? 1.0f : 0.0f;
    }
  for (int i = 0; i < 4; ++i) {
    if (aug[i][i] == 0.0f) {
      int pick = -1;
      float best = 0.0f;
      for (int j = i + 1; j < 4; ++j)
        if (absf(aug[j][i]) > best) { best = absf(aug[j][i]); pick = j; }
      if (pick < 0) return false;
      for (int c = 0; c < 8; ++c) { const float t = aug[i][c]; aug[i][c] = aug[pick][c]; aug[pick][c] = t; }
    }
    const float scale = 1.0f / aug[i][i];
    for (int c = 0; c < 8; ++c) aug[i][c] *= scale;
    for (int r = 0; r < 4; ++r) {
      if (r == i) continue;
      const float f = aug[r][i];
      if (f == 0.0f) continue;
      for (int c = 0; c < 8; ++c) aug[r][c] -= aug[i][c] * f;
    }
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) m[4 * r + c] = aug[r][4 + c];
  return true;
}

// VolumetricData::cell_axes, VolumetricData.C:138-168
VT_HD void cell_axes(const vt_grid &g, float xd[3], float yd[3], float zd[3]) {
  const float rx = g.xsize > 1 ? 1.0f / (g.xsize - 1.0f) : 1.0f;
  const float ry = g.ysize > 1 ? 1.0f / (g.ysize - 1.0f) : 1.0f;
  const float rz = g.zsize > 1 ? 1.0f / (g.zsize - 1.0f) : 1.0f;
  for (int k = 0; k < 3; ++k) {
    xd[k] = (float)(g.xaxis[k] * rx);
    yd[k] = (float)(g.yaxis[k] * ry);
    zd[k] = (float)(g.zaxis[k] * rz);
  }
}

// voxel_coord_from_cartesian_coord's matrix (VolumetricData.C:194-225), inverted once
VT_HD void grid_inverse(const vt_grid &g, bool shift, float inv[16]) {
  const int nx1 = g.xsize - 1, ny1 = g.ysize - 1, nz1 = g.zsize - 1;
  float corner[3];
  for (int k = 0; k < 3; ++k) {
    corner[k] = (float)g.origin[k];
    if (shift) corner[k] = (float)g.origin[k] - 0.5f * (float)(g.xaxis[k] / nx1 + g.yaxis[k] / ny1 + g.zaxis[k] / nz1);
  }
  for (int k = 0; k < 3; ++k) {
    inv[4 * k + 0] = (float)g.xaxis[k] / nx1;
    inv[4 * k + 1] = (float)g.yaxis[k] / ny1;
    inv[4 * k + 2] = (float)g.zaxis[k] / nz1;
    inv[4 * k + 3] = 0.f;
  }
  inv[12] = corner[0]; inv[13] = corner[1]; inv[14] = corner[2]; inv[15] = 1.f;
  invert4(inv);
}

VT_HD double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// samples along one output axis: the denser of the two source samplings, n/L (not (n-1)/L),
// truncated -- Voltool.C:69-74 / 138-143
VT_HD int axis_samples(const double out[3], const double a[3], int na, const double b[3], int nb) {
  const double viaA = dot3(out, a) * na / dot3(a, a);
  const double viaB = dot3(out, b) * nb / dot3(b, b);
  const double m = maxd(viaA, viaB);
  // C's (int) of a NaN / out-of-range double is INT_MIN on x86-64
  return (m >= -2147483648.0 && m < 2147483648.0) ? (int)m : (int)0x80000000;
}

// init_from_intersection, Voltool.C:51-78
VT_HD void intersection_grid(const vt_grid &A, const vt_grid &B, vt_grid &o) {
  for (int d = 0; d < 3; ++d) {
    const double lo = maxd(A.origin[d], B.origin[d]);
    o.origin[d] = lo;
    o.xaxis[d] = maxd(mind(A.origin[d] + A.xaxis[d], B.origin[d] + B.xaxis[d]), lo);
    o.yaxis[d] = maxd(mind(A.origin[d] + A.yaxis[d], B.origin[d] + B.yaxis[d]), lo);
    o.zaxis[d] = maxd(mind(A.origin[d] + A.zaxis[d], B.origin[d] + B.zaxis[d]), lo);
  }
  for (int d = 0; d < 3; ++d) {
    o.xaxis[d] -= o.origin[d];
    o.yaxis[d] -= o.origin[d];
    o.zaxis[d] -= o.origin[d];
  }
  o.xsize = axis_samples(o.xaxis, A.xaxis, A.xsize, B.xaxis, B.xsize);
  o.ysize = axis_samples(o.yaxis, A.yaxis, A.ysize, B.yaxis, B.ysize);
  o.zsize = axis_samples(o.zaxis, A.zaxis, A.zsize, B.zaxis, B.zsize);
}

// init_from_union, Voltool.C:114-147 (diagonal only: orthogonal cells)
VT_HD void union_grid(const vt_grid &A, const vt_grid &B, vt_grid &o) {
  for (int d = 0; d < 3; ++d) {
    o.origin[d] = mind(A.origin[d], B.origin[d]);
    o.xaxis[d] = o.yaxis[d] = o.zaxis[d] = 0.0;
  }
  o.xaxis[0] = maxd(maxd(A.origin[0] + A.xaxis[0], B.origin[0] + B.xaxis[0]), o.origin[0]) - o.origin[0];
  o.yaxis[1] = maxd(maxd(A.origin[1] + A.yaxis[1], B.origin[1] + B.yaxis[1]), o.origin[1]) - o.origin[1];
  o.zaxis[2] = maxd(maxd(A.origin[2] + A.zaxis[2], B.origin[2] + B.zaxis[2]), o.origin[2]) - o.origin[2];
  o.xsize = axis_samples(o.xaxis, A.xaxis, A.xsize, B.xaxis, B.xsize);
  o.ysize = axis_samples(o.yaxis, A.yaxis, A.ysize, B.yaxis, B.ysize);
  o.zsize = axis_samples(o.zaxis, A.zaxis, A.zsize, B.zaxis, B.zsize);
}

// ---- separable lookup tables -------------------------------------------------------------------
// For orthogonal axis-aligned cells every off-diagonal term of voxel_coord -> multpoint3d is an
// exact zero, so output index g along one axis fixes the fractional source coordinate along that
// axis.  One entry caches what voxel_value_interpolate needs from it.
struct
#if defined(__CUDACC__)
    __align__(16)
#endif
    AxisEntry {
  int i0, i1;  // neighbour indices after voxel_value_safe's clamp (VolumetricData.C:253-260)
  float f;     // fraction (VolumetricData.C:265-270)
  int inside;  // truncated index within [0, n): the bounds test of :315-321
};

// out_origin/out_delta: origin and cell delta of the OUTPUT grid along this axis
// inv_diag/inv_trans:   inverse-matrix terms of the SOURCE map along this axis
VT_HD AxisEntry axis_entry(int axis, int g, double out_origin, float out_delta, float inv_diag, float inv_trans,
                           int n_src) {
  // voxel_coord: foreign-axis products are index*0 and add nothing; kept in order
  const float fg = (float)g;
  float prod[3] = {0.f, 0.f, 0.f};
  prod[axis] = fg * out_delta;
  const float cart = (float)(((out_origin + (double)prod[0]) + (double)prod[1]) + (double)prod[2]);
  // multpoint3d with w == 1
  float t[3] = {0.f, 0.f, 0.f};
  t[axis] = cart * inv_diag;
  const float vox = ((t[0] + t[1]) + t[2]) + inv_trans;
  const int iv = (vox >= -2147483648.0f && vox < 2147483648.0f) ? (int)vox : (int)0x80000000;
  AxisEntry e;
  e.inside = (iv >= 0 && iv < n_src) ? 1 : 0;
  e.f = vox - (float)iv;
  e.i0 = iv > 0 ? (iv < n_src ? iv : n_src - 1) : 0;
  const long long i1 = (long long)iv + 1;
  e.i1 = i1 > 0 ? (i1 < n_src ? (int)i1 : n_src - 1) : 0;
  return e;
}

VT_HD bool diagonal_inverse(const float *m) {
  return m[1] == 0.f && m[2] == 0.f && m[4] == 0.f && m[6] == 0.f && m[8] == 0.f && m[9] == 0.f && m[3] == 0.f &&
         m[7] == 0.f && m[11] == 0.f && m[15] == 1.0f;
}

}  // namespace geom
}  // namespace vt
