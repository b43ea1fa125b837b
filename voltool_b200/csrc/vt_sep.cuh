// vt_sep.cuh -- table-driven trilinear sampling for the separable (axis-aligned) path, shared by the
// two-map CC kernel and the batched fit CC kernel.
//
// A warp owns 32 consecutive output x, kRows consecutive output y and a chunk of output z.  For each
// source map a thread keeps the bilinear (x then y) value of the two source planes it is between;
// stepping z by one usually turns the upper plane into the lower one, so only one new plane is
// sampled per step.  Within a plane, source rows shared by neighbouring output rows are x-lerped
// once.  Every value is produced by the reference's expression
// (voxel_value_interpolate, VolumetricData.C:264-291): a0 + xf*(a1-a0), then y, then z.
#pragma once
#include "vt_device.cuh"
#include "vt_geom.h"

namespace vt {
using geom::AxisEntry;

constexpr int kRows = 4;

template <int Y>
struct MapMarch {
  const float *__restrict__ d;
  long pstride;
  // per thread (x)
  int x0, x1;
  float xf;
  unsigned int loadmask;  // bit 2r: row r0[r] must be loaded, bit 2r+1: row r1[r] (else reuse)
  // per output row (warp-uniform)
  int r0[Y], r1[Y];   // row offsets inside a plane (< 2^31 voxels per plane)
  float yf[Y];
  // plane cache
  int z0, z1;
  float lo[Y], hi[Y];
};

template <int Y>
__device__ __forceinline__ void march_init(MapMarch<Y> &m, const float *d, int nx_src, int ny_src, const AxisEntry &ex,
                                           const AxisEntry *__restrict__ ty, int gy0, int ny_out) {
  m.d = d;
  m.pstride = (long)nx_src * ny_src;
  m.x0 = ex.i0; m.x1 = ex.i1; m.xf = ex.f;
#pragma unroll
  for (int r = 0; r < Y; ++r) {
    const int gy = min(gy0 + r, ny_out - 1);
    const AxisEntry ey = ty[gy];
    m.r0[r] = ey.i0 * nx_src;
    m.r1[r] = ey.i1 * nx_src;
    m.yf[r] = ey.f;
  }
  m.loadmask = 0u;
#pragma unroll
  for (int r = 0; r < Y; ++r) {
    if (r == 0 || m.r0[r] != m.r1[r - 1]) m.loadmask |= 1u << (2 * r);
    if (m.r1[r] != m.r0[r]) m.loadmask |= 1u << (2 * r + 1);
  }
  m.z0 = -1; m.z1 = -1;
}

// Bilinear values of plane z for the Y rows.  Both x neighbours are plain loads (they share a
// sector almost always; L1 serves the second): a shuffle would save the request but costs a
// WARPSYNC/collective pair per row in SASS.  All loads of the plane are issued before the first
// lerp; which rows need loading is warp-uniform (m.loadmask).
template <int Y>
__device__ __forceinline__ void march_plane(const MapMarch<Y> &m, int z, float (&out)[Y]) {
  const float *__restrict__ P0 = m.d + (long)z * m.pstride + m.x0;
  const float *__restrict__ P1 = m.d + (long)z * m.pstride + m.x1;
  float a0[Y], a1[Y], b0[Y], b1[Y];
#pragma unroll
  for (int r = 0; r < Y; ++r) {
    a0[r] = a1[r] = b0[r] = b1[r] = 0.f;
    if ((m.loadmask >> (2 * r)) & 1) { a0[r] = __ldg(P0 + m.r0[r]); a1[r] = __ldg(P1 + m.r0[r]); }
    if ((m.loadmask >> (2 * r + 1)) & 1) { b0[r] = __ldg(P0 + m.r1[r]); b1[r] = __ldg(P1 + m.r1[r]); }
  }
  float prev_xl = 0.f;
#pragma unroll
  for (int r = 0; r < Y; ++r) {
    // same source row as the previous output row's upper row -> reuse its x lerp
    const float xa = ((m.loadmask >> (2 * r)) & 1) ? a0[r] + m.xf * (a1[r] - a0[r]) : prev_xl;
    // clamped edge: both rows are the same source row
    const float xb = ((m.loadmask >> (2 * r + 1)) & 1) ? b0[r] + m.xf * (b1[r] - b0[r]) : xa;
    prev_xl = xb;
    out[r] = xa + m.yf[r] * (xb - xa);
  }
}

// values of the Y rows at the output z whose table entry is ez (ez.inside must hold)
template <int Y>
__device__ __forceinline__ void march_sample(MapMarch<Y> &m, const AxisEntry &ez, float (&v)[Y]) {
  if (ez.i0 == m.z1 && ez.i0 != m.z0) {  // the usual advance: upper plane becomes the lower one
#pragma unroll
    for (int r = 0; r < Y; ++r) { const float t = m.lo[r]; m.lo[r] = m.hi[r]; m.hi[r] = t; }
    const int t = m.z0; m.z0 = m.z1; m.z1 = t;
  }
  if (ez.i0 != m.z0) { march_plane(m, ez.i0, m.lo); m.z0 = ez.i0; }
  if (ez.i1 == m.z0) {
#pragma unroll
    for (int r = 0; r < Y; ++r) m.hi[r] = m.lo[r];
    m.z1 = m.z0;
  } else if (ez.i1 != m.z1) {
    march_plane(m, ez.i1, m.hi);
    m.z1 = ez.i1;
  }
#pragma unroll
  for (int r = 0; r < Y; ++r) v[r] = m.lo[r] + ez.f * (m.hi[r] - m.lo[r]);
}

// smallest float T with (double)v >= thr  <=>  v >= T for every float v (MDFF.C:77 compares the
// float sample against the double threshold)
inline float threshold_as_float(double thr) {
  if (thr != thr) return __builtin_nanf("");
  float t = (float)thr;                      // round to nearest
  if ((double)t < thr) t = __builtin_nextafterf(t, __builtin_inff());
  return t;
}

struct CcAcc {
  double s[5];
  long long n;
};

// One warp work item: 32 x, kRows y, output z in [zbeg, zend).  Tables: tA*/tB* per axis.
// Adds this thread's voxels to its double partial sums.
template <int Y>
__device__ __forceinline__ void cc_march_item(const float *__restrict__ A, int anx, int any, const float *__restrict__ B,
                                              int bnx, int bny, const AxisEntry *__restrict__ tAx,
                                              const AxisEntry *__restrict__ tAy, const AxisEntry *__restrict__ tAz,
                                              const AxisEntry *__restrict__ tBx, const AxisEntry *__restrict__ tBy,
                                              const AxisEntry *__restrict__ tBz, int nx, int ny, int gx, int gy0,
                                              int zbeg, int zend, float thr_f, CcAcc &acc) {
  const int gxc = min(gx, nx - 1);
  const AxisEntry ax = tAx[gxc], bx = tBx[gxc];
  MapMarch<Y> ma, mb;
  march_init(ma, A, anx, any, ax, tAy, gy0, ny);
  march_init(mb, B, bnx, bny, bx, tBy, gy0, ny);
  bool ok[Y];
  bool any_ok = false;
#pragma unroll
  for (int r = 0; r < Y; ++r) {
    const int gy = gy0 + r;
    const int gyc = min(gy, ny - 1);
    ok[r] = gx < nx && gy < ny && ax.inside && bx.inside && tAy[gyc].inside && tBy[gyc].inside;
    any_ok |= ok[r];
  }
  if (!__any_sync(0xffffffffu, any_ok)) return;
  for (int gz = zbeg; gz < zend; ++gz) {
    const AxisEntry az = tAz[gz], bz = tBz[gz];
    if (!az.inside || !bz.inside) continue;  // warp-uniform
    float va[Y], vb[Y];
    march_sample(ma, az, va);
    bool want[Y];
    bool any_want = false;
#pragma unroll
    for (int r = 0; r < Y; ++r) {
      want[r] = ok[r] && (va[r] == va[r]) && (va[r] >= thr_f);   // !isnan(vA) && vA >= thr (MDFF.C:73-77)
      any_want |= want[r];
    }
    if (!__any_sync(0xffffffffu, any_want)) continue;            // B is only needed where A passes
    march_sample(mb, bz, vb);
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < Y; ++r) {
      const bool take = want[r] && (vb[r] == vb[r]);
      if (take) {  // MDFF.C:78-83: double sums of float terms (float products)
        acc.s[0] += (double)va[r]; acc.s[1] += (double)vb[r];
        acc.s[2] += (double)(va[r] * va[r]); acc.s[3] += (double)(vb[r] * vb[r]); acc.s[4] += (double)(va[r] * vb[r]);
        ++cnt;
      }
    }
    acc.n += cnt;
  }
}

}  // namespace vt
