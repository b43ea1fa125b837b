// vt_host.cpp -- host-side metadata arithmetic (see vt_host.h).  Build: g++ -O2 -ffp-contract=off.
#include "vt_host.h"
#include "vt_geom.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace vt {

static const double kPi = 3.14159265358979323846;  // VMD_PI, utilities.h:41

// ------------------------------------------------------------------------------------------
// Matrix4 (Matrix4.C is not part of the snapshot; definition shared with oracle/, "parity
// unpinned").  Float arithmetic throughout.
// ------------------------------------------------------------------------------------------
void mat4_identity(Mat4 &a) {
  std::memset(a.m, 0, sizeof(a.m));
  a.m[0] = a.m[5] = a.m[10] = a.m[15] = 1.0f;
}

bool mat4_invert(Mat4 &M) { return geom::invert4(M.m); }

// projective transform: scale by 1/w first, then the affine part
void mat4_point(const Mat4 &A, const float in[3], float out[3]) {
  const float *m = A.m;
  const float iw = 1.0f / (in[0] * m[3] + in[1] * m[7] + in[2] * m[11] + m[15]);
  const float u = iw * in[0], v = iw * in[1], w = iw * in[2];
  const float rx = u * m[0] + v * m[4] + w * m[8] + iw * m[12];
  const float ry = u * m[1] + v * m[5] + w * m[9] + iw * m[13];
  const float rz = u * m[2] + v * m[6] + w * m[10] + iw * m[14];
  out[0] = rx; out[1] = ry; out[2] = rz;
}

void mat4_axis_rotation(Mat4 &A, int axis, float angle_rad) {
  mat4_identity(A);
  const float c = (float)std::cos((double)angle_rad);
  const float s = (float)std::sin((double)angle_rad);
  float *m = A.m;
  switch (axis) {
    case 0: m[5] = c; m[6] = s; m[9] = -s; m[10] = c; break;
    case 1: m[0] = c; m[2] = -s; m[8] = s; m[10] = c; break;
    default: m[0] = c; m[1] = s; m[4] = -s; m[5] = c; break;
  }
}

// ------------------------------------------------------------------------------------------
// grid geometry
// ------------------------------------------------------------------------------------------
long gridsize(const vt_grid &g) { return (long)g.xsize * (long)g.ysize * (long)g.zsize; }

void cell_axes(const vt_grid &g, float xd[3], float yd[3], float zd[3]) { geom::cell_axes(g, xd, yd, zd); }
void grid_inverse(const vt_grid &g, bool shift, float inv[16]) { geom::grid_inverse(g, shift, inv); }

void map_geom(const vt_grid &g, MapGeom &o) {
  grid_inverse(g, false, o.inv0);
  grid_inverse(g, true, o.inv1);
  o.nx = g.xsize; o.ny = g.ysize; o.nz = g.zsize;
}

void coord_geom(const vt_grid &g, CoordGeom &o) {
  for (int k = 0; k < 3; ++k) o.origin[k] = g.origin[k];
  cell_axes(g, o.xd, o.yd, o.zd);
  o.nx = g.xsize; o.ny = g.ysize; o.nz = g.zsize;
}

void intersection_grid(const vt_grid &A, const vt_grid &B, vt_grid &o) { geom::intersection_grid(A, B, o); }
void union_grid(const vt_grid &A, const vt_grid &B, vt_grid &o) { geom::union_grid(A, B, o); }

void pad_plan(vt_grid &g, int pxm, int pxp, int pym, int pyp, int pzm, int pzp, PadPlan &p) {
  const int ox = g.xsize, oy = g.ysize, oz = g.zsize;
  double dx[3], dy[3], dz[3];
  for (int k = 0; k < 3; ++k) {
    dx[k] = g.xaxis[k] / (ox - 1);
    dy[k] = g.yaxis[k] / (oy - 1);
    dz[k] = g.zaxis[k] / (oz - 1);
  }
  p.nx = std::max(1, ox + pxm + pxp);
  p.ny = std::max(1, oy + pym + pyp);
  p.nz = std::max(1, oz + pzm + pzp);
  p.sx = std::max(0, pxm); p.sy = std::max(0, pym); p.sz = std::max(0, pzm);
  p.ex = std::min(p.nx, ox + pxm); p.ey = std::min(p.ny, oy + pym); p.ez = std::min(p.nz, oz + pzm);
  p.pxm = pxm; p.pym = pym; p.pzm = pzm;
  g.xsize = p.nx; g.ysize = p.ny; g.zsize = p.nz;
  // vec_scaled_add(double*, float, const double*) utilities.h:222-226: the int count becomes a float
  const float gx = (float)(pxm + pxp), gy = (float)(pym + pyp), gz = (float)(pzm + pzp);
  const float sx = (float)(-pxm), sy = (float)(-pym), sz = (float)(-pzm);
  for (int k = 0; k < 3; ++k) g.xaxis[k] += gx * dx[k];
  for (int k = 0; k < 3; ++k) g.yaxis[k] += gy * dy[k];
  for (int k = 0; k < 3; ++k) g.zaxis[k] += gz * dz[k];
  for (int k = 0; k < 3; ++k) g.origin[k] += sx * dx[k];
  for (int k = 0; k < 3; ++k) g.origin[k] += sy * dy[k];
  for (int k = 0; k < 3; ++k) g.origin[k] += sz * dz[k];
}

void crop_pads(const vt_grid &g, double minx, double miny, double minz, double maxx, double maxy, double maxz,
               int pads[6]) {
  const double hx = g.xaxis[0] / (g.xsize - 1);
  const double hy = g.yaxis[1] / (g.ysize - 1);
  const double hz = g.zaxis[2] / (g.zsize - 1);
  pads[0] = int((g.origin[0] - minx) / hx);
  pads[2] = int((g.origin[1] - miny) / hy);
  pads[4] = int((g.origin[2] - minz) / hz);
  pads[1] = int((maxx - g.origin[0] - g.xaxis[0]) / hx);
  pads[3] = int((maxy - g.origin[1] - g.yaxis[1]) / hy);
  pads[5] = int((maxz - g.origin[2] - g.zaxis[2]) / hz);
}

void downsample_geom(vt_grid &g) {
  g.xsize /= 2; g.ysize /= 2; g.zsize /= 2;
  // the scale is formed from the NEW sizes with an integer division (VolumetricData.C:750-752)
  const double fx = 0.5 * (g.xsize) / (g.xsize / 2);
  const double fy = 0.5 * (g.ysize) / (g.ysize / 2);
  const double fz = 0.5 * (g.zsize) / (g.zsize / 2);
  for (int k = 0; k < 3; ++k) { g.xaxis[k] *= fx; g.yaxis[k] *= fy; g.zaxis[k] *= fz; }
}

void supersample_geom(vt_grid &g) {
  g.xsize = g.xsize * 2 - 1; g.ysize = g.ysize * 2 - 1; g.zsize = g.zsize * 2 - 1;  // axes unchanged
}

// vol_moveto: translate the map so that `com` lands on `pos`; the origin is carried in float
void vol_moveto(vt_grid &g, const float com[3], const float pos[3]) {
  for (int k = 0; k < 3; ++k) {
    const float o = (float)g.origin[k];
    const float shift = pos[k] - com[k];
    const float moved = o + shift;
    g.origin[k] = moved;
  }
}

// vol_move: rigid transform of the map frame by a column-major 4x4; origin in float, axes via
// vectrans (double * float summed in double, stored through float)
void vol_move(vt_grid &g, const float m[16]) {
  const float o[3] = {(float)g.origin[0], (float)g.origin[1], (float)g.origin[2]};
  for (int k = 0; k < 3; ++k) {
    const float rot = o[0] * m[k] + o[1] * m[4 + k] + o[2] * m[8 + k];
    const float moved = rot + m[12 + k];
    g.origin[k] = moved;
  }
  double *axes[3] = {g.xaxis, g.yaxis, g.zaxis};
  for (int a = 0; a < 3; ++a) {
    const double v[3] = {axes[a][0], axes[a][1], axes[a][2]};
    for (int k = 0; k < 3; ++k) axes[a][k] = (float)(v[0] * m[k] + v[1] * m[4 + k] + v[2] * m[8 + k]);
  }
}

double cc_from_sums(const double s[5], long n) {
  const int size = (int)n;
  const double mA = s[0] / size, mB = s[1] / size;
  const double sdA = std::sqrt(s[2] / size - mA * mA);
  const double sdB = std::sqrt(s[3] / size - mB * mB);
  return (s[4] - size * mA * mB) / (size * sdA * sdB);
}

// GaussianBlur (Segmentation.h / GaussianBlur.C not in the snapshot): radius ceil(3 sigma) >= 1,
// weights expf(-k^2/(2 sigma^2)) divided by their float sum taken left to right.
int blur_taps(double sigma, float *w, int maxw) {
  int R = (int)std::ceil(3.0 * sigma);
  if (R < 1) R = 1;
  if (2 * R + 1 > maxw) return -1;
  const float sf = (float)sigma;
  const float two_var = 2.0f * sf * sf;
  float total = 0.0f;
  for (int k = -R; k <= R; ++k) {
    const float v = expf(-(float)(k * k) / two_var);
    w[k + R] = v;
    total += v;
  }
  for (int k = 0; k <= 2 * R; ++k) w[k] = w[k] / total;
  return R;
}

// QuickSurf bounding-box padding (restatement shared with oracle/): max(p, 0.65*sqrt(4/3*pi*p^3)),
// p = radscale*maxrad*1.70
float sim_padding(float radscale, float maxrad) {
  float pad = radscale * maxrad * 1.70f;
  float padrad = pad;
  padrad = 0.65f * sqrtf(4.0f / 3.0f * ((float)kPi) * padrad * padrad * padrad);
  return std::max(pad, padrad);
}

float gausslim_for_quality(int q) { return q == 3 ? 4.0f : q == 2 ? 3.0f : q == 1 ? 2.5f : 2.0f; }

void sim_policy(double resolution, double gspacing_in, float &radscale, double &gspacing, int &quality) {
  radscale = (float)(.2 * resolution);                 // TclMDFF.C:125, TclVoltool.C:81
  gspacing = gspacing_in;
  if (gspacing == 0) gspacing = 1.5 * radscale;        // TclMDFF.C:492-494, TclVoltool.C:82
  quality = resolution >= 9 ? 0 : 3;                   // TclMDFF.C:130-135
}

void euler_cs(double degrees, float &c, float &s) {
  const double amount = degrees * kPi / 180.0;         // DEGTORAD(stride*amt)
  const float a = (float)amount;                       // rotate_axis takes a float
  c = (float)std::cos((double)a);
  s = (float)std::sin((double)a);
}

void pose_apply(float *pos, long n, const float com[3], const vt_pose &pose) {
  const float neg[3] = {-1.0f * com[0], -1.0f * com[1], -1.0f * com[2]};  // TclVoltool.C:157
  Mat4 R[3];
  for (int ax = 0; ax < 3; ++ax) {
    mat4_identity(R[ax]);
    float c, s;
    euler_cs((double)pose.rot_deg[ax], c, s);
    float *m = R[ax].m;
    if (ax == 0) { m[5] = c; m[6] = s; m[9] = -s; m[10] = c; }
    else if (ax == 1) { m[0] = c; m[2] = -s; m[8] = s; m[10] = c; }
    else { m[0] = c; m[1] = s; m[4] = -s; m[5] = c; }
  }
  for (long i = 0; i < n; ++i) {
    float *p = pos + 3 * i;
    p[0] = p[0] + neg[0]; p[1] = p[1] + neg[1]; p[2] = p[2] + neg[2];
    mat4_point(R[0], p, p);
    mat4_point(R[1], p, p);
    mat4_point(R[2], p, p);
    p[0] = p[0] + com[0]; p[1] = p[1] + com[1]; p[2] = p[2] + com[2];
    p[0] = p[0] + pose.shift[0]; p[1] = p[1] + pose.shift[1]; p[2] = p[2] + pose.shift[2];
  }
}

// measure_center (Measure.C not in the snapshot): weighted mean, float products summed in double
void measure_center(const float *pos, const float *w, long n, float com[3]) {
  double sx = 0, sy = 0, sz = 0, sw = 0;
  for (long i = 0; i < n; ++i) {
    const float wi = w[i];
    sw += (double)wi;
    sx += (double)(wi * pos[3 * i]);
    sy += (double)(wi * pos[3 * i + 1]);
    sz += (double)(wi * pos[3 * i + 2]);
  }
  com[0] = (float)(sx / sw); com[1] = (float)(sy / sw); com[2] = (float)(sz / sw);
}

// rotate()'s candidate loop with the order-dependent skip (TclVoltool.C:160-199): a candidate
// that scores below the running best also jumps over the next z.
int rotate_replay(const float *table, int m, float *returncc, int *best_idx) {
  int visited = 0;
  *best_idx = -1;
  for (int x = 0; x < m; ++x)
    for (int y = 0; y < m; ++y)
      for (int z = 0; z < m; ++z) {
        const int idx = (x * m + y) * m + z;
        const float cc = table[idx];
        ++visited;
        if (cc > *returncc) { *returncc = cc; *best_idx = idx; }
        else if (cc < *returncc) ++z;
      }
  return visited;
}

}  // namespace vt
