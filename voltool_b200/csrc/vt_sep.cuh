// vt_sep.cuh -- table-driven trilinear sampling for the separable (axis-aligned) path, shared by the
// two-map CC kernel and the batched fit CC kernel.
//
// A warp owns 32 consecutive output x, kRows consecutive output y and a chunk of output z.  For each
// source map a thread keeps the bilinear (x then y) value of the two source planes it is between;
// stepping z by one usually turns the upper plane into the lower one, so only one new plane is
// sampled per step.  Every value is produced by the reference's expression
// (voxel_value_interpolate, VolumetricData.C:264-291): a0 + xf*(a1-a0), then y, then z.
#pragma once
#include "vt_device.cuh"
#include "vt_geom.h"

namespace vt {
using geom::AxisEntry;

constexpr int kRows = 4;

// a + f * (b - a) on a packed pair with the reference's three roundings (sub, mul, add): FADD2, FMUL2 and
// an FFMA2 whose multiplier is 1.0 (= an exactly rounded add; written that way because ptxas would
// otherwise contract mul.rn.f32x2 + add.rn.f32x2 into one fused FFMA2, which rounds once)
// (the 1.0 pair comes from constant memory: a literal would let ptxas fold the FFMA2 back into an add
// and then contract it with the multiply after all)
static __constant__ f32x2 c_one_pair = 0x3f8000003f800000ull;
__device__ __forceinline__ f32x2 lerp2(f32x2 a, f32x2 b, f32x2 f, f32x2 one) {
  return fma2(mul2(f, sub2(b, a)), one, a);
}

template <int Y>
struct MapMarch {
  static_assert(Y % 2 == 0, "rows are processed as packed pairs");
  const char *__restrict__ base;
  size_t pstride;          // bytes per source plane
  // per thread (x)
  unsigned int xoff, dx;   // byte offset of the left x neighbour inside a row; right neighbour = + dx (0: clamped)
  f32x2 xf;
  // per output row (warp-uniform): byte offsets of the two source rows inside a plane (< 2^30 voxels per plane)
  unsigned int r0[Y], r1[Y];
  f32x2 yf[Y / 2];
  f32x2 one;               // (1.0, 1.0) from constant memory, see lerp2
  // plane cache
  int z0, z1;
  f32x2 lo[Y / 2], hi[Y / 2];
};

template <int Y>
__device__ __forceinline__ void march_init(MapMarch<Y> &m, const float *d, int nx_src, int ny_src, const AxisEntry &ex,
                                           const AxisEntry *__restrict__ ty, int gy0, int ny_out) {
  m.base = reinterpret_cast<const char *>(d);
  m.pstride = (size_t)nx_src * ny_src * sizeof(float);
  m.xoff = (unsigned int)ex.i0 * 4u;
  m.dx = (unsigned int)(ex.i1 - ex.i0) * 4u;
  m.xf = pack2(ex.f, ex.f);
  m.one = c_one_pair;
  float yf[Y];
#pragma unroll
  for (int r = 0; r < Y; ++r) {
    const int gy = min(gy0 + r, ny_out - 1);
    const AxisEntry ey = ty[gy];
    m.r0[r] = (unsigned int)ey.i0 * (unsigned int)nx_src * 4u;
    m.r1[r] = (unsigned int)ey.i1 * (unsigned int)nx_src * 4u;
    yf[r] = ey.f;
  }
#pragma unroll
  for (int q = 0; q < Y / 2; ++q) m.yf[q] = pack2(yf[2 * q], yf[2 * q + 1]);
  m.z0 = -1; m.z1 = -1;
}

// Bilinear values of plane z for the Y rows, two output rows per packed value.  Every output row
// loads its four neighbours (no branches; rows shared between neighbouring output rows and the
// right-hand x neighbour are L1 hits), all loads are issued before the first lerp.
template <int Y>
__device__ __forceinline__ void march_plane(const MapMarch<Y> &m, int z, f32x2 (&out)[Y / 2]) {
  const char *__restrict__ P0 = m.base + (size_t)z * m.pstride + m.xoff;
  const char *__restrict__ P1 = P0 + m.dx;
  float a0[Y], a1[Y], b0[Y], b1[Y];
#pragma unroll
  for (int r = 0; r < Y; ++r) {
    a0[r] = __ldg(reinterpret_cast<const float *>(P0 + m.r0[r]));
    a1[r] = __ldg(reinterpret_cast<const float *>(P1 + m.r0[r]));
    b0[r] = __ldg(reinterpret_cast<const float *>(P0 + m.r1[r]));
    b1[r] = __ldg(reinterpret_cast<const float *>(P1 + m.r1[r]));
  }
#pragma unroll
  for (int q = 0; q < Y / 2; ++q) {
    const f32x2 xa = lerp2(pack2(a0[2 * q], a0[2 * q + 1]), pack2(a1[2 * q], a1[2 * q + 1]), m.xf, m.one);
    const f32x2 xb = lerp2(pack2(b0[2 * q], b0[2 * q + 1]), pack2(b1[2 * q], b1[2 * q + 1]), m.xf, m.one);
    out[q] = lerp2(xa, xb, m.yf[q], m.one);
  }
}

// values of the Y rows at the output z whose table entry is ez (ez.inside must hold)
template <int Y>
__device__ __forceinline__ void march_sample(MapMarch<Y> &m, const AxisEntry &ez, float (&v)[Y]) {
  if (ez.i0 == m.z1 && ez.i0 != m.z0) {  // the usual advance: upper plane becomes the lower one
#pragma unroll
    for (int q = 0; q < Y / 2; ++q) { const f32x2 t = m.lo[q]; m.lo[q] = m.hi[q]; m.hi[q] = t; }
    const int t = m.z0; m.z0 = m.z1; m.z1 = t;
  }
  if (ez.i0 != m.z0) { march_plane(m, ez.i0, m.lo); m.z0 = ez.i0; }
  if (ez.i1 == m.z0) {
#pragma unroll
    for (int q = 0; q < Y / 2; ++q) m.hi[q] = m.lo[q];
    m.z1 = m.z0;
  } else if (ez.i1 != m.z1) {
    march_plane(m, ez.i1, m.hi);
    m.z1 = ez.i1;
  }
  const f32x2 zf = pack2(ez.f, ez.f);
#pragma unroll
  for (int q = 0; q < Y / 2; ++q) unpack2(lerp2(m.lo[q], m.hi[q], zf, m.one), v[2 * q], v[2 * q + 1]);
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
