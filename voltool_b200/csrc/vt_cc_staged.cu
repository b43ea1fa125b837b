// vt_cc_staged.cu -- HBM-streaming variant of the separable CC kernel for large maps.
//
// The direct kernel (vt_cc.cu / vt_sep.cuh) loads source rows straight into registers; with ~1 us of
// loaded HBM latency and a dependent chain per z step it cannot keep enough bytes in flight.  Here
// the loads are decoupled from the arithmetic: two producer warps stream whole source-plane boxes
// (rows of the tile + halo) into a shared-memory ring with cp.async.bulk (the TMA copy engine,
// completion on mbarriers), several planes ahead; eight consumer warps take planes from the ring,
// form the bilinear value of their rows once per plane and march in z exactly like the direct
// kernel.  Same table-driven arithmetic (bit-identical samples); the five sums are formed in float
// over the Y rows of one z step before they are added in double, so CC agrees with the direct path
// to ~1e-8 rather than 1e-12.  Selected for large grids only (or VOLTOOL_CC_STAGED=1).
#include <cstdlib>
#include <vector>

#include "vt_device.cuh"
#include "vt_geom.h"
#include "vt_sep.cuh"
#include "vt_tma.h"

namespace vt {
namespace staged {

constexpr int TX = 64;       // output x per CTA tile (2 warps x 32 lanes)
constexpr int XS = 72;       // floats per staged source row (TX + 1 halo + 16-byte alignment slack)
constexpr int S = 4;         // ring stages per map
constexpr int kConsumers = 8;
constexpr int kThreadsStaged = (kConsumers + 2) * 32;

template <int Y> struct Shape {
  static constexpr int TY = 4 * Y;     // 4 warp rows x Y rows per thread
  static constexpr int YS = TY + 4;    // staged source rows per plane (TY + 1 halo + slack)
};

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// producers are far ahead of the consumers most of the time: back off instead of burning issue slots
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long *bar, unsigned int parity) {
  unsigned int done = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(200);
  }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned int bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct Box {  // source box of one tile in one map
  int xb, wfl;   // first staged x (multiple of 4), floats per row (multiple of 4)
  int yb, nrows;
};

__device__ __forceinline__ Box tile_box(const AxisEntry *__restrict__ tx, const AxisEntry *__restrict__ ty, int gx0,
                                        int gx1, int gy0, int gy1, int src_nx) {
  Box b;
  b.xb = tx[gx0].i0 & ~3;
  int w = (tx[gx1].i1 - b.xb + 1 + 3) & ~3;
  if (b.xb + w > src_nx) w = src_nx - b.xb;
  b.wfl = w;
  b.yb = ty[gy0].i0;
  b.nrows = ty[gy1].i1 - b.yb + 1;
  return b;
}

template <int Y>
struct Consumer {
  // unit-stride fast path: every output row/column advances the source index by exactly one
  int base;      // (first source row - box row 0) * XS + (x0 - box x 0)
  int base1;     // same for x1 (x1 - x0 is 0 at a clamped edge, else 1)
  bool unit;     // warp-uniform
  // generic path re-reads the y table per plane
  const AxisEntry *ty;
  int gy0, ny_out, yb, x0l, x1l;
  float xf;
  float yf[Y];
  int z0, z1;
  float lo[Y], hi[Y];
};

template <int Y>
__device__ __forceinline__ void consumer_init(Consumer<Y> &c, const AxisEntry &ex, const AxisEntry *__restrict__ ty,
                                              int gy0, int ny_out, const Box &b) {
  c.x0l = ex.i0 - b.xb; c.x1l = ex.i1 - b.xb; c.xf = ex.f;
  c.ty = ty; c.gy0 = gy0; c.ny_out = ny_out; c.yb = b.yb;
  bool rows_unit = true;
  int first = 0, prev = 0;
#pragma unroll
  for (int r = 0; r < Y; ++r) {
    const AxisEntry ey = ty[min(gy0 + r, ny_out - 1)];
    c.yf[r] = ey.f;
    if (r == 0) first = ey.i0;
    else rows_unit = rows_unit && (ey.i0 == prev + 1);
    rows_unit = rows_unit && (ey.i1 == ey.i0 + 1);
    prev = ey.i0;
  }
  c.unit = rows_unit;
  c.base = (first - b.yb) * XS + c.x0l;
  c.base1 = (first - b.yb) * XS + c.x1l;
  c.z0 = -1; c.z1 = -1;
}

// bilinear values of the staged plane for this thread's rows: a0 + xf*(a1-a0) per source row,
// then the y lerp (voxel_value_interpolate, VolumetricData.C:273-287)
template <int Y, bool KNOWN_UNIT = false>
__device__ __forceinline__ void consumer_plane(const Consumer<Y> &c, const float *__restrict__ P, float (&out)[Y]) {
  if (KNOWN_UNIT || c.unit) {
    const float *__restrict__ q = P + c.base;
    const float *__restrict__ q1 = P + c.base1;
    float xl[Y + 1];
#pragma unroll
    for (int r = 0; r <= Y; ++r) {
      const float v0 = q[r * XS], v1 = q1[r * XS];
      xl[r] = v0 + c.xf * (v1 - v0);
    }
#pragma unroll
    for (int r = 0; r < Y; ++r) out[r] = xl[r] + c.yf[r] * (xl[r + 1] - xl[r]);
    return;
  }
  int prev_row = -1;
  float prev_xl = 0.f;
#pragma unroll
  for (int r = 0; r < Y; ++r) {
    const AxisEntry ey = c.ty[min(c.gy0 + r, c.ny_out - 1)];
    const int ra = (ey.i0 - c.yb) * XS, rb = (ey.i1 - c.yb) * XS;
    float xa, xb;
    if (ra == prev_row) xa = prev_xl;
    else { const float v0 = P[ra + c.x0l], v1 = P[ra + c.x1l]; xa = v0 + c.xf * (v1 - v0); }
    if (rb == ra) xb = xa;
    else { const float v0 = P[rb + c.x0l], v1 = P[rb + c.x1l]; xb = v0 + c.xf * (v1 - v0); }
    prev_row = rb;
    prev_xl = xb;
    out[r] = xa + ey.f * (xb - xa);
  }
}

// take the next plane of this map's ring (planes arrive strictly in order)
template <int Y, bool KNOWN_UNIT = false>
__device__ __forceinline__ void consumer_take(const Consumer<Y> &c, const float *__restrict__ ring,
                                              unsigned long long *full, unsigned long long *empty, unsigned int &cnt,
                                              int plane_floats, float (&out)[Y]) {
  const unsigned int slot = cnt % S, par = (cnt / S) & 1u;
  mbar_wait(&full[slot], par);
  consumer_plane<Y, KNOWN_UNIT>(c, ring + (size_t)slot * plane_floats, out);
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[slot]);
  ++cnt;
}

template <int Y>
__device__ __forceinline__ void consumer_sample(Consumer<Y> &c, const AxisEntry &ez, const float *__restrict__ ring,
                                                unsigned long long *full, unsigned long long *empty, unsigned int &cnt,
                                                int plane_floats, float (&v)[Y]) {
  if (ez.i0 == c.z1 && ez.i0 != c.z0) {
#pragma unroll
    for (int r = 0; r < Y; ++r) { const float t = c.lo[r]; c.lo[r] = c.hi[r]; c.hi[r] = t; }
    const int t = c.z0; c.z0 = c.z1; c.z1 = t;
  }
  if (ez.i0 != c.z0) { consumer_take<Y>(c, ring, full, empty, cnt, plane_floats, c.lo); c.z0 = ez.i0; }
  if (ez.i1 == c.z0) {
#pragma unroll
    for (int r = 0; r < Y; ++r) c.hi[r] = c.lo[r];
    c.z1 = c.z0;
  } else if (ez.i1 != c.z1) {
    consumer_take<Y>(c, ring, full, empty, cnt, plane_floats, c.hi);
    c.z1 = ez.i1;
  }
#pragma unroll
  for (int r = 0; r < Y; ++r) v[r] = c.lo[r] + ez.f * (c.hi[r] - c.lo[r]);
}

__device__ __forceinline__ void producer_run(const float *__restrict__ src, int snx, int sny, const Box &b, int p0,
                                             int p1, float *ring, unsigned long long *full, unsigned long long *empty,
                                             unsigned int &cnt, int plane_floats) {
  const int lane = threadIdx.x & 31;
  const unsigned int wbytes = (unsigned int)b.wfl * 4u;
  for (int p = p0; p <= p1; ++p) {
    const unsigned int slot = cnt % S, par = ((cnt / S) & 1u) ^ 1u;
    mbar_wait_backoff(&empty[slot], par);
    if (lane == 0) mbar_expect_tx(&full[slot], wbytes * (unsigned int)b.nrows);
    __syncwarp();
    float *dst = ring + (size_t)slot * plane_floats;
    const float *g = src + ((size_t)p * sny + b.yb) * snx + b.xb;
    for (int r = lane; r < b.nrows; r += 32) bulk_g2s(dst + r * XS, g + (size_t)r * snx, wbytes, &full[slot]);
    ++cnt;
  }
}

// every step of [za, zb) uses source planes (k, k+1) with k advancing by exactly one (warp-wide check)
__device__ __forceinline__ bool chunk_regular(const AxisEntry *__restrict__ t, int za, int zb) {
  bool ok = true;
  for (int g = za + (int)(threadIdx.x & 31); g < zb; g += 32) {
    const AxisEntry e = t[g];
    ok = ok && (e.i1 == e.i0 + 1);
    if (g + 1 < zb) ok = ok && (t[g + 1].i0 == e.i0 + 1);
  }
  return __all_sync(0xffffffffu, ok);
}

// MODE 0: cross correlation sums.  MODE 1..5: two-map op written to outp (Voltool.C:249-377):
// 1 add, 2 subtract, 3 multiply (intersection rule), 4 multiply (union NaN rule), 5 average.
struct BinaryArgs {
  float *outp;
  DevMap gA, gB;   // for the exact re-evaluation of voxels whose result is a NaN (payload rules)
  DevCoord gC;
};

template <int Y, int MODE>
__global__ void __launch_bounds__(kThreadsStaged, (Y > 4 ? 1 : 2))
    cc_staged_kernel(const float *__restrict__ A, int anx, int any, const float *__restrict__ B, int bnx, int bny,
                     const AxisEntry *__restrict__ tAx, const AxisEntry *__restrict__ tAy,
                     const AxisEntry *__restrict__ tAz, const AxisEntry *__restrict__ tBx,
                     const AxisEntry *__restrict__ tBy, const AxisEntry *__restrict__ tBz, int nx, int ny, int nz,
                     int zchunk, int zlo, int zhi, float thr_f, double *partials, unsigned int *ticket, double *out,
                     BinaryArgs ba) {
  constexpr int TY = Shape<Y>::TY, YS = Shape<Y>::YS;
  constexpr int plane_floats = YS * XS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *ringA = reinterpret_cast<float *>(smem_raw);
  float *ringB = ringA + (size_t)S * plane_floats;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ringB + (size_t)S * plane_floats);
  unsigned long long *fullA = bars, *emptyA = bars + S, *fullB = bars + 2 * S, *emptyB = bars + 3 * S;
  __shared__ double sm[6 * 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&fullA[s], 1); mbar_init(&fullB[s], 1);
      mbar_init(&emptyA[s], kConsumers); mbar_init(&emptyB[s], kConsumers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  double acc[6] = {0, 0, 0, 0, 0, 0};
  long long nvox = 0;
  unsigned int cntA = 0, cntB = 0;
  const int xtiles = (nx + TX - 1) / TX, ytiles = (ny + TY - 1) / TY;
  const int zchunks = (zhi - zlo + zchunk - 1) / zchunk;
  const long work = (long)xtiles * ytiles * zchunks;
  for (long w = blockIdx.x; w < work; w += gridDim.x) {
    const int xt = (int)(w % xtiles);
    const long rr = w / xtiles;
    const int yt = (int)(rr % ytiles), zc = (int)(rr / ytiles);
    const int gx_first = xt * TX, gx_last = min(nx - 1, gx_first + TX - 1);
    const int gy_first = yt * TY, gy_last = min(ny - 1, gy_first + TY - 1);
    const int za = zlo + zc * zchunk, zb = min(zhi, za + zchunk);  // every z in [zlo, zhi) is inside both maps
    const Box bA = tile_box(tAx, tAy, gx_first, gx_last, gy_first, gy_last, anx);
    const Box bB = tile_box(tBx, tBy, gx_first, gx_last, gy_first, gy_last, bnx);
    if (warp == kConsumers) {
      producer_run(A, anx, any, bA, tAz[za].i0, tAz[zb - 1].i1, ringA, fullA, emptyA, cntA, plane_floats);
    } else if (warp == kConsumers + 1) {
      producer_run(B, bnx, bny, bB, tBz[za].i0, tBz[zb - 1].i1, ringB, fullB, emptyB, cntB, plane_floats);
    } else {
      const int gx = gx_first + (warp & 1) * 32 + lane;
      const int gy0 = gy_first + (warp >> 1) * Y;
      const int gxc = min(gx, nx - 1);
      const AxisEntry ax = tAx[gxc], bx = tBx[gxc];
      Consumer<Y> ca, cb;
      consumer_init<Y>(ca, ax, tAy, gy0, ny, bA);
      consumer_init<Y>(cb, bx, tBy, gy0, ny, bB);
      bool ok[Y], inA[Y], inB[Y];
#pragma unroll
      for (int r = 0; r < Y; ++r) {
        const int gy = gy0 + r, gyc = min(gy, ny - 1);
        inA[r] = ax.inside && tAy[gyc].inside;
        inB[r] = bx.inside && tBy[gyc].inside;
        ok[r] = gx < nx && gy < ny;
      }
      // one output plane of this warp's rows: the CC sums or the two-map op
      auto emit = [&](int gz, const float (&va)[Y], const float (&vb)[Y]) {
          if (MODE == 0) {
            float fA = 0.f, fB = 0.f, fAA = 0.f, fBB = 0.f, fAB = 0.f;
            int cnt = 0;
  #pragma unroll
            for (int r = 0; r < Y; ++r) {
              const bool take = ok[r] && inA[r] && inB[r] && (va[r] >= thr_f) && (vb[r] == vb[r]);  // MDFF.C:73-77
              const float a = take ? va[r] : 0.f, b = take ? vb[r] : 0.f;
              fA += a; fB += b;
              fAA += a * a; fBB += b * b; fAB += a * b;
              cnt += take ? 1 : 0;
            }
            acc[0] += (double)fA; acc[1] += (double)fB; acc[2] += (double)fAA; acc[3] += (double)fBB;
            acc[4] += (double)fAB;
            nvox += cnt;
          } else {
            constexpr bool SAFE = (MODE != 3 && MODE != 4);  // add/sub/avg read 0 outside, multiply reads NaN
  #pragma unroll
            for (int r = 0; r < Y; ++r) {
              if (!ok[r]) continue;
              const int gy = gy0 + r;
              float a = inA[r] ? va[r] : (SAFE ? 0.0f : quiet_nan());
              float b = inB[r] ? vb[r] : (SAFE ? 0.0f : quiet_nan());
              if (a != a || b != b) {
                // a NaN is involved: redo this voxel with the literal per-voxel arithmetic so that the
                // x86 payload rules of the reference are reproduced
                float x, y, z;
                voxel_coord(ba.gC, gx, gy, gz, x, y, z);
                a = sample_interp<SAFE>(ba.gA, x, y, z);
                b = sample_interp<SAFE>(ba.gB, x, y, z);
              }
              float v;
              if (MODE == 1) v = nan_like_x86(a + b, a, b);
              else if (MODE == 2) v = nan_like_x86(a - b, a, b);
              else if (MODE == 3) v = nan_like_x86(a * b, a, b);
              else if (MODE == 4) {
                const bool na = is_nan_bits(a), nb = is_nan_bits(b);
                v = (!na && !nb) ? nan_like_x86(a * b, a, b) : (!na ? a : (!nb ? b : 0.0f));
              } else v = nan_like_x86((float)((double)(a + b) * 0.5), a, b);
              ba.outp[((size_t)gz * ny + gy) * nx + gx] = v;
            }
          }
      };
      // Regular chunk (the usual case: both maps advance one source plane per output plane, rows
      // and columns unit-stride): after the first two planes every z step takes exactly one new
      // plane per map, and the two plane buffers swap roles by unrolling -- no table-driven
      // branching, no register moves.  Same arithmetic as the general loop below.
      const bool lean = ca.unit && cb.unit && chunk_regular(tAz, za, zb) && chunk_regular(tBz, za, zb);
      if (lean) {
        float pa[Y], qa[Y], pb[Y], qb[Y], va[Y], vb[Y];
        consumer_take<Y, true>(ca, ringA, fullA, emptyA, cntA, plane_floats, pa);
        consumer_take<Y, true>(cb, ringB, fullB, emptyB, cntB, plane_floats, pb);
        auto step = [&](int gz, const float (&la)[Y], float (&ha)[Y], const float (&lb)[Y], float (&hb)[Y]) {
          const float fa = tAz[gz].f, fb = tBz[gz].f;
          consumer_take<Y, true>(ca, ringA, fullA, emptyA, cntA, plane_floats, ha);
          consumer_take<Y, true>(cb, ringB, fullB, emptyB, cntB, plane_floats, hb);
#pragma unroll
          for (int r = 0; r < Y; ++r) {
            va[r] = la[r] + fa * (ha[r] - la[r]);
            vb[r] = lb[r] + fb * (hb[r] - lb[r]);
          }
          emit(gz, va, vb);
        };
        int gz = za;
        for (; gz + 1 < zb; gz += 2) {
          step(gz, pa, qa, pb, qb);
          step(gz + 1, qa, pa, qb, pb);
        }
        if (gz < zb) step(gz, pa, qa, pb, qb);
      } else {
        for (int gz = za; gz < zb; ++gz) {
          const AxisEntry az = tAz[gz], bz = tBz[gz];
          float va[Y], vb[Y];
          consumer_sample<Y>(ca, az, ringA, fullA, emptyA, cntA, plane_floats, va);
          consumer_sample<Y>(cb, bz, ringB, fullB, emptyB, cntB, plane_floats, vb);
          emit(gz, va, vb);
        }
      }
    }
  }
  if (MODE == 0) {
    acc[5] = (double)nvox;
    block_sum<6>(acc, sm);
    grid_finish<6>(acc, partials, ticket, out, sm);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// The packed form for REGULAR sampling: every output row / plane advances the source row / plane of both maps by
// exactly one (the upper neighbour clamped at the map's last row / plane) -- what two maps of equal spacing give,
// whatever their offset.  The host verifies that from the tables (lean_ok) and hands over only the first source
// row / plane of each map.  A lane owns TWO output columns, x and x + 32, and carries every value as a packed pair
// over them: the x, y and z lerps of voxel_value_interpolate (VolumetricData.C:273-287) are then three packed
// instructions per pair with the reference's three roundings each (lerp2, vt_sep.cuh), the planes swap roles by
// unrolling, and the CC terms are packed products and packed adds.  Half the arithmetic instructions per voxel of
// the scalar ring kernel above, the same ring, the same samples bit for bit.
struct LeanMap {
  const float *d;
  int nx, ny, nz;    // source map
  int y0, z0;        // source row of output row 0, source plane of output plane zlo
};

// shape (measured at 512^3 / 1024^3, correlate with a half-voxel offset): 6 warps x 4 rows at 128 registers 0.396 / 2.84 ms;
// 8 warps x 2 rows at 94 registers 0.355 / 2.71 ms (kept: more warps hide the ring and table latencies better than
// longer rows amortise the x lerps); 10 warps x 2 rows at 80 registers (spills) 0.438 / 3.16 ms; six ring stages
// instead of four: no change
#ifndef LEAN_WARPS
#define LEAN_WARPS 8
#define LEAN_Y 2
#define LEAN_CTAS 2
#endif
constexpr int kLeanWarps = LEAN_WARPS;                           // consumer warps, one group of LEAN_Y rows each
constexpr int kLeanThreads = (kLeanWarps + 2) * 32;

template <int Y> struct LeanShape {
  static constexpr int TY = kLeanWarps * Y;
  static constexpr int YS = TY + 4;                     // staged rows per plane: TY + the upper neighbour, rounded so that a
                                                        // plane is a multiple of 128 bytes (TMA destinations)
  static_assert((YS * XS * 4) % 128 == 0, "ring stages must stay 128-byte aligned");
};

struct LeanCols {   // one map, this lane's two columns
  int a0, a1, b0, b1;   // staged x of the lower / upper neighbour of column x (a) and x + 32 (b)
  f32x2 fx;             // their x fractions
};

struct LeanRows {   // staged rows of this warp's Y output rows + the upper neighbour of the last: consecutive, except that
  int first;        // a row past the end of the staged box (the clamped upper neighbour, rows below a ragged tile edge)
  int last;         // reads the box's last row
};

template <int Y>
__device__ __forceinline__ void lean_plane(const float *__restrict__ P, const LeanCols &c, const LeanRows &rows,
                                           const f32x2 (&yf)[Y], f32x2 one, f32x2 (&out)[Y]) {
  f32x2 xl[Y + 1];
#pragma unroll
  for (int r = 0; r <= Y; ++r) {
    const float *__restrict__ q = P + min(rows.first + r, rows.last) * XS;
    const f32x2 v0 = pack2(q[c.a0], q[c.b0]), v1 = pack2(q[c.a1], q[c.b1]);
    xl[r] = lerp2(v0, v1, c.fx, one);
  }
#pragma unroll
  for (int r = 0; r < Y; ++r) out[r] = lerp2(xl[r], xl[r + 1], yf[r], one);
}

template <int Y>
__device__ __forceinline__ void lean_take(const float *__restrict__ ring, unsigned long long *full, unsigned long long *empty,
                                          unsigned int &cnt, int plane_floats, const LeanCols &c, const LeanRows &rowoff,
                                          const f32x2 (&yf)[Y], f32x2 one, f32x2 (&out)[Y]) {
  const unsigned int slot = cnt % S, par = (cnt / S) & 1u;
  mbar_wait(&full[slot], par);
  lean_plane<Y>(ring + (size_t)slot * plane_floats, c, rowoff, yf, one, out);
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[slot]);
  ++cnt;
}

template <int Y, int MODE>
__global__ void __launch_bounds__(kLeanThreads, LEAN_CTAS)
    cc_lean_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, LeanMap A, LeanMap B,
                   const AxisEntry *__restrict__ tAx, const AxisEntry *__restrict__ tAy,
                   const AxisEntry *__restrict__ tAz, const AxisEntry *__restrict__ tBx, const AxisEntry *__restrict__ tBy,
                   const AxisEntry *__restrict__ tBz, int nx, int ny, int zchunk, int zlo, int zhi, float thr_f,
                   double *partials, unsigned int *ticket, double *out, BinaryArgs ba) {
  constexpr int TY = LeanShape<Y>::TY, YS = LeanShape<Y>::YS;
  constexpr int plane_floats = YS * XS;
  constexpr int LTX = 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *ringA = reinterpret_cast<float *>(smem_raw);
  float *ringB = ringA + (size_t)S * plane_floats;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ringB + (size_t)S * plane_floats);
  unsigned long long *fullA = bars, *emptyA = bars + S, *fullB = bars + 2 * S, *emptyB = bars + 3 * S;
  __shared__ double sm[6 * 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&fullA[s], 1); mbar_init(&fullB[s], 1);
      mbar_init(&emptyA[s], kLeanWarps); mbar_init(&emptyB[s], kLeanWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  double acc[6] = {0, 0, 0, 0, 0, 0};
  long long nvox = 0;
  unsigned int cntA = 0, cntB = 0;
  const f32x2 one = c_one_pair;
  const int xtiles = (nx + LTX - 1) / LTX, ytiles = (ny + TY - 1) / TY;
  const int zchunks = (zhi - zlo + zchunk - 1) / zchunk;
  const long work = (long)xtiles * ytiles * zchunks;
  for (long w = blockIdx.x; w < work; w += gridDim.x) {
    const int xt = (int)(w % xtiles);
    const long rr = w / xtiles;
    const int yt = (int)(rr % ytiles), zc = (int)(rr / ytiles);
    const int gx_first = xt * LTX, gx_last = min(nx - 1, gx_first + LTX - 1);
    const int gy_first = yt * TY, gy_last = min(ny - 1, gy_first + TY - 1);
    const int za = zlo + zc * zchunk, zb = min(zhi, za + zchunk);
    // staged boxes: x from the tables (16-byte aligned start), rows y0 + gy_first .. upper neighbour of the last row
    const LeanMap *maps[2] = {&A, &B};
    int xbm[2], wflm[2], ybm[2], nrowsm[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const AxisEntry *tx = m ? tBx : tAx;
      xbm[m] = tx[gx_first].i0 & ~3;
      int wd = (tx[gx_last].i1 - xbm[m] + 1 + 3) & ~3;
      if (xbm[m] + wd > maps[m]->nx) wd = maps[m]->nx - xbm[m];
      wflm[m] = wd;
      ybm[m] = maps[m]->y0 + gy_first;
      nrowsm[m] = min(maps[m]->y0 + gy_last + 1, maps[m]->ny - 1) - ybm[m] + 1;
    }
    if (warp >= kLeanWarps) {   // producers, one per map: ONE tensor-map box load (72 x YS x 1) per source plane
      const int m = warp - kLeanWarps;
      const LeanMap &M = *maps[m];
      const int p0 = M.z0 + (za - zlo), p1 = min(M.z0 + (zb - 1 - zlo) + 1, M.nz - 1);
      float *ring = m ? ringB : ringA;
      unsigned long long *full = m ? fullB : fullA, *empty = m ? emptyB : emptyA;
      unsigned int &cnt = m ? cntB : cntA;
      for (int p = p0; p <= p1; ++p) {
        const unsigned int slot = cnt % S, par = ((cnt / S) & 1u) ^ 1u;
        mbar_wait_backoff(&empty[slot], par);
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(&full[slot], (unsigned int)(plane_floats * sizeof(float)));
          tma_load_box3d(smem_u32(ring + (size_t)slot * plane_floats), m ? &tmB : &tmA, xbm[m], ybm[m], p, smem_u32(&full[slot]));
        }
        __syncwarp();
        ++cnt;
      }
      continue;
    }
    const int gxa = gx_first + lane, gxb = gxa + 32;
    const int gy0 = gy_first + warp * Y;
    LeanCols ca, cb;
    bool colin[2];   // column inside both maps and inside the output grid
    bool colA[2], colB[2];
    {
      const AxisEntry aa = tAx[min(gxa, nx - 1)], ab = tAx[min(gxb, nx - 1)];
      const AxisEntry ba_ = tBx[min(gxa, nx - 1)], bb = tBx[min(gxb, nx - 1)];
      ca.a0 = aa.i0 - xbm[0]; ca.a1 = aa.i1 - xbm[0]; ca.b0 = ab.i0 - xbm[0]; ca.b1 = ab.i1 - xbm[0];
      cb.a0 = ba_.i0 - xbm[1]; cb.a1 = ba_.i1 - xbm[1]; cb.b0 = bb.i0 - xbm[1]; cb.b1 = bb.i1 - xbm[1];
      ca.fx = pack2(aa.f, ab.f); cb.fx = pack2(ba_.f, bb.f);
      colin[0] = gxa < nx && aa.inside && ba_.inside;
      colin[1] = gxb < nx && ab.inside && bb.inside;
      colA[0] = aa.inside; colA[1] = ab.inside; colB[0] = ba_.inside; colB[1] = bb.inside;
    }
    LeanRows rowA, rowB;
    f32x2 yfA[Y], yfB[Y];
    bool rowin[Y], rowA_in[Y], rowB_in[Y];
    rowA.first = rowB.first = gy0 - gy_first;   // staged row of output row gy0 + r: first + r
    rowA.last = nrowsm[0] - 1; rowB.last = nrowsm[1] - 1;
#pragma unroll
    for (int r = 0; r < Y; ++r) {
      const int gyc = min(gy0 + r, ny - 1);
      const AxisEntry ea = tAy[gyc], eb = tBy[gyc];
      yfA[r] = pack2(ea.f, ea.f); yfB[r] = pack2(eb.f, eb.f);
      rowin[r] = gy0 + r < ny && ea.inside && eb.inside;
      rowA_in[r] = ea.inside; rowB_in[r] = eb.inside;
    }

    auto emit = [&](int gz, const f32x2 (&va)[Y], const f32x2 (&vb)[Y]) {
      if (MODE == 0) {
        f32x2 fA = pack2(0.f, 0.f), fB = fA, fAA = fA, fBB = fA, fAB = fA;
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < Y; ++r) {
          float a0, a1, b0, b1;
          unpack2(va[r], a0, a1);
          unpack2(vb[r], b0, b1);
          const bool t0 = rowin[r] && colin[0] && (a0 >= thr_f) && (b0 == b0);   // MDFF.C:73-77
          const bool t1 = rowin[r] && colin[1] && (a1 >= thr_f) && (b1 == b1);
          const f32x2 a = pack2(t0 ? a0 : 0.f, t1 ? a1 : 0.f), b = pack2(t0 ? b0 : 0.f, t1 ? b1 : 0.f);
          fA = add2(fA, a); fB = add2(fB, b);
          fAA = fma2(mul2(a, a), one, fAA);   // float product, then an exactly rounded add (MDFF.C:81-83)
          fBB = fma2(mul2(b, b), one, fBB);
          fAB = fma2(mul2(a, b), one, fAB);
          cnt += (t0 ? 1 : 0) + (t1 ? 1 : 0);
        }
        float x, y;
        unpack2(fA, x, y); acc[0] += (double)x; acc[0] += (double)y;
        unpack2(fB, x, y); acc[1] += (double)x; acc[1] += (double)y;
        unpack2(fAA, x, y); acc[2] += (double)x; acc[2] += (double)y;
        unpack2(fBB, x, y); acc[3] += (double)x; acc[3] += (double)y;
        unpack2(fAB, x, y); acc[4] += (double)x; acc[4] += (double)y;
        nvox += cnt;
      } else {
        constexpr bool SAFE = (MODE != 3 && MODE != 4);  // add/sub/avg read 0 outside, multiply reads NaN
        auto op = [](float a, float b) -> float {
          if (MODE == 1) return nan_like_x86(a + b, a, b);
          if (MODE == 2) return nan_like_x86(a - b, a, b);
          if (MODE == 3) return nan_like_x86(a * b, a, b);
          if (MODE == 4) {
            const bool na = is_nan_bits(a), nb = is_nan_bits(b);
            return (!na && !nb) ? nan_like_x86(a * b, a, b) : (!na ? a : (!nb ? b : 0.0f));
          }
          return nan_like_x86((float)((double)(a + b) * 0.5), a, b);
        };
        float v[Y][2];
        unsigned int redo = 0;   // voxels with a NaN operand: redone below with the literal per-voxel arithmetic
#pragma unroll
        for (int r = 0; r < Y; ++r) {
          float av[2], bv[2];
          unpack2(va[r], av[0], av[1]);
          unpack2(vb[r], bv[0], bv[1]);
#pragma unroll
          for (int cidx = 0; cidx < 2; ++cidx) {
            const float a = (colA[cidx] && rowA_in[r]) ? av[cidx] : (SAFE ? 0.0f : quiet_nan());
            const float b = (colB[cidx] && rowB_in[r]) ? bv[cidx] : (SAFE ? 0.0f : quiet_nan());
            if (a != a || b != b) redo |= 1u << (2 * r + cidx);
            v[r][cidx] = op(a, b);
          }
        }
#pragma unroll
        for (int r = 0; r < Y; ++r) {
          const int gy = gy0 + r;
          if (gy >= ny) continue;
          float *__restrict__ orow = ba.outp + ((size_t)gz * ny + gy) * nx;
          if (gxa < nx) orow[gxa] = v[r][0];
          if (gxb < nx) orow[gxb] = v[r][1];
        }
        if (redo) {   // rare: the x86 payload rules of the reference need the operands as the literal chain forms them
#pragma unroll 1
          for (int k = 0; k < 2 * Y; ++k) {
            if (!((redo >> k) & 1u)) continue;
            const int gy = gy0 + (k >> 1), gx = (k & 1) ? gxb : gxa;
            if (gy >= ny || gx >= nx) continue;
            float x, y, z;
            voxel_coord(ba.gC, gx, gy, gz, x, y, z);
            const float a = sample_interp<SAFE>(ba.gA, x, y, z), b = sample_interp<SAFE>(ba.gB, x, y, z);
            ba.outp[((size_t)gz * ny + gy) * nx + gx] = op(a, b);
          }
        }
      }
    };

    f32x2 pa[Y], qa[Y], pb[Y], qb[Y], va[Y], vb[Y];
    lean_take<Y>(ringA, fullA, emptyA, cntA, plane_floats, ca, rowA, yfA, one, pa);
    lean_take<Y>(ringB, fullB, emptyB, cntB, plane_floats, cb, rowB, yfB, one, pb);
    // the upper plane of step gz is source plane z0 + (gz - zlo) + 1, clamped at the map's last plane: only the very
    // last step of a map that ends with the output grid can be clamped, and then upper = lower
    auto step = [&](int gz, const f32x2 (&la)[Y], f32x2 (&ha)[Y], const f32x2 (&lb)[Y], f32x2 (&hb)[Y]) {
      const float fa = tAz[gz].f, fb = tBz[gz].f;
      const f32x2 fa2 = pack2(fa, fa), fb2 = pack2(fb, fb);
      if (A.z0 + (gz - zlo) + 1 <= A.nz - 1) lean_take<Y>(ringA, fullA, emptyA, cntA, plane_floats, ca, rowA, yfA, one, ha);
      else {
#pragma unroll
        for (int r = 0; r < Y; ++r) ha[r] = la[r];
      }
      if (B.z0 + (gz - zlo) + 1 <= B.nz - 1) lean_take<Y>(ringB, fullB, emptyB, cntB, plane_floats, cb, rowB, yfB, one, hb);
      else {
#pragma unroll
        for (int r = 0; r < Y; ++r) hb[r] = lb[r];
      }
#pragma unroll
      for (int r = 0; r < Y; ++r) {
        va[r] = lerp2(la[r], ha[r], fa2, one);
        vb[r] = lerp2(lb[r], hb[r], fb2, one);
      }
      emit(gz, va, vb);
    };
    int gz = za;
    for (; gz + 1 < zb; gz += 2) {
      step(gz, pa, qa, pb, qb);
      step(gz + 1, qa, pa, qb, pb);
    }
    if (gz < zb) step(gz, pa, qa, pb, qb);
  }
  if (MODE == 0) {
    acc[5] = (double)nvox;
    block_sum<6>(acc, sm);
    grid_finish<6>(acc, partials, ticket, out, sm);
  }
}

}  // namespace staged

// Host: is the staged kernel applicable to these tables?  `tab` = [Ax|Ay|Az|Bx|By|Bz] with offsets.
template <int Y>
static bool staged_ok(const std::vector<AxisEntry> &tab, const int off[6], int nx, int ny, int nz, int anx, int bnx,
                      int &zlo, int &zhi) {
  using namespace staged;
  constexpr int TY = Shape<Y>::TY, YS = Shape<Y>::YS;
  if ((anx & 3) || (bnx & 3) || nx < 1 || ny < 1 || nz < 1) return false;
  const int srcnx[2] = {anx, bnx};
  for (int m = 0; m < 2; ++m) {
    const AxisEntry *tx = &tab[off[3 * m]], *ty = &tab[off[3 * m + 1]], *tz = &tab[off[3 * m + 2]];
    for (int g = 1; g < nx; ++g) if (tx[g].i0 < tx[g - 1].i0 || tx[g].i1 < tx[g - 1].i1) return false;
    for (int g = 1; g < ny; ++g) if (ty[g].i0 < ty[g - 1].i0 || ty[g].i1 < ty[g - 1].i1) return false;
    for (int g0 = 0; g0 < nx; g0 += TX) {
      const int g1 = std::min(nx - 1, g0 + TX - 1);
      const int xb = tx[g0].i0 & ~3;
      int w = (tx[g1].i1 - xb + 1 + 3) & ~3;
      if (xb + w > srcnx[m]) w = srcnx[m] - xb;
      if (w > XS || w < 1 || tx[g1].i1 - xb >= w) return false;
    }
    for (int g0 = 0; g0 < ny; g0 += TY) {
      const int g1 = std::min(ny - 1, g0 + TY - 1);
      if (ty[g1].i1 - ty[g0].i0 + 1 > YS) return false;
    }
    (void)tz;
  }
  // z: the steps inside both maps must form one interval, and the planes they touch must be
  // consecutive in each map (a step may advance the lower plane by at most its own upper plane + 1)
  const AxisEntry *az = &tab[off[2]], *bz = &tab[off[5]];
  zlo = -1; zhi = -1;
  for (int g = 0; g < nz; ++g) {
    const bool in = az[g].inside && bz[g].inside;
    if (in) { if (zlo < 0) zlo = g; if (zhi >= 0 && zhi != g) return false; zhi = g + 1; }
  }
  if (zlo < 0) return false;
  for (int m = 0; m < 2; ++m) {
    const AxisEntry *tz = m ? bz : az;
    for (int g = zlo; g < zhi; ++g) {
      if (tz[g].i1 != tz[g].i0 && tz[g].i1 != tz[g].i0 + 1) return false;
      if (g > zlo && (tz[g].i0 < tz[g - 1].i0 || tz[g].i0 > tz[g - 1].i1 + 1)) return false;
    }
  }
  return true;
}

// Host: regular sampling (see cc_lean_kernel)?  Both maps: every output row / plane inside [0, ny) x [zlo, zhi) steps
// the source row / plane by one, upper neighbour = lower + 1 clamped at the last; x monotone and narrow enough per
// 64-wide tile.
template <int Y>
static bool lean_ok(const std::vector<AxisEntry> &tab, const int off[6], int nx, int ny, int zlo, int zhi, const DevMap &A,
                    const DevMap &B, staged::LeanMap lm[2]) {
  using namespace staged;
  const DevMap *maps[2] = {&A, &B};
  for (int m = 0; m < 2; ++m) {
    const DevMap &M = *maps[m];
    if (M.nx & 3) return false;
    const AxisEntry *tx = &tab[off[3 * m]], *ty = &tab[off[3 * m + 1]], *tz = &tab[off[3 * m + 2]];
    for (int g = 0; g < ny; ++g) {
      if (!ty[g].inside || ty[g].i0 != ty[0].i0 + g || ty[g].i1 != std::min(ty[g].i0 + 1, M.ny - 1)) return false;
    }
    for (int g = zlo; g < zhi; ++g) {
      if (!tz[g].inside || tz[g].i0 != tz[zlo].i0 + (g - zlo) || tz[g].i1 != std::min(tz[g].i0 + 1, M.nz - 1)) return false;
    }
    for (int g = 1; g < nx; ++g) if (tx[g].i0 < tx[g - 1].i0 || tx[g].i1 < tx[g - 1].i1) return false;
    for (int g0 = 0; g0 < nx; g0 += 64) {
      const int g1 = std::min(nx - 1, g0 + 63);
      const int xb = tx[g0].i0 & ~3;
      int w = (tx[g1].i1 - xb + 1 + 3) & ~3;
      if (xb + w > M.nx) w = M.nx - xb;
      if (w > XS || w < 1 || tx[g1].i1 - xb >= w) return false;
    }
    lm[m].d = M.d; lm[m].nx = M.nx; lm[m].ny = M.ny; lm[m].nz = M.nz;
    lm[m].y0 = ty[0].i0; lm[m].z0 = tz[zlo].i0;
  }
  return true;
}

template <int Y, int MODE>
static int launch_lean(const staged::LeanMap lm[2], int nx, int ny, int zlo, int zhi, const int off[6], const AxisEntry *d_tab,
                       float thr_f, double *partials, unsigned int *ticket, double *dout, const staged::BinaryArgs &ba) {
  using namespace staged;
  Ctx &c = ctx();
  constexpr int TY = LeanShape<Y>::TY, YS = LeanShape<Y>::YS;
  const size_t smem = sizeof(float) * 2 * (size_t)S * YS * XS + sizeof(unsigned long long) * 4 * S;
  VT_CUDA(cudaFuncSetAttribute(cc_lean_kernel<Y, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUtensorMap tmA, tmB;
  if (tma_encode_map3d(&tmA, lm[0].d, lm[0].nx, lm[0].ny, lm[0].nz, XS, YS, 1, false) ||
      tma_encode_map3d(&tmB, lm[1].d, lm[1].nx, lm[1].ny, lm[1].nz, XS, YS, 1, false))
    return 1;   // no tensor maps: the caller takes the row-copy ring kernel
  const int zchunk = 64;
  const long work = (long)((nx + 63) / 64) * ((ny + TY - 1) / TY) * ((zhi - zlo + zchunk - 1) / zchunk);
  const int grid = (int)std::min<long>(work, (long)c.sm_count * LEAN_CTAS);
  prof_begin(MODE == 0 ? PK_CC : PK_BINARY);
  cc_lean_kernel<Y, MODE><<<grid, kLeanThreads, smem, c.stream>>>(tmA, tmB, lm[0], lm[1], d_tab + off[0], d_tab + off[1], d_tab + off[2],
                                                                  d_tab + off[3], d_tab + off[4], d_tab + off[5], nx, ny, zchunk,
                                                                  zlo, zhi, thr_f, partials, ticket, dout, ba);
  prof_end(MODE == 0 ? PK_CC : PK_BINARY);
  VT_LAUNCHED();
  return 0;
}

template <int Y, int MODE>
static int launch_staged(const DevMap &A, const DevMap &B, int nx, int ny, int nz, const std::vector<AxisEntry> &host_tab,
                         const int off[6], const AxisEntry *d_tab, float thr_f, double *partials, unsigned int *ticket,
                         double *dout, const staged::BinaryArgs &ba, int *zlo_out, int *zhi_out) {
  using namespace staged;
  int zlo, zhi;
  if (!staged_ok<Y>(host_tab, off, nx, ny, nz, A.nx, B.nx, zlo, zhi)) return 1;
  if (zlo_out) { *zlo_out = zlo; *zhi_out = zhi; }
  {
    LeanMap lm[2];
    if (!getenv("VOLTOOL_CC_NOLEAN") && lean_ok<LEAN_Y>(host_tab, off, nx, ny, zlo, zhi, A, B, lm)) {
      const int lrc = launch_lean<LEAN_Y, MODE>(lm, nx, ny, zlo, zhi, off, d_tab, thr_f, partials, ticket, dout, ba);
      if (lrc != 1) return lrc;
    }
  }
  Ctx &c = ctx();
  constexpr int TY = Shape<Y>::TY, YS = Shape<Y>::YS;
  const size_t smem = sizeof(float) * 2 * (size_t)S * YS * XS + sizeof(unsigned long long) * 4 * S;
  VT_CUDA(cudaFuncSetAttribute(cc_staged_kernel<Y, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int zchunk = 64;
  const long work = (long)((nx + TX - 1) / TX) * ((ny + TY - 1) / TY) * ((zhi - zlo + zchunk - 1) / zchunk);
  int grid = (int)std::min<long>(work, (long)c.sm_count * (Y > 4 ? 1 : 2));
  prof_begin(MODE == 0 ? PK_CC : PK_BINARY);
  cc_staged_kernel<Y, MODE><<<grid, kThreadsStaged, smem, c.stream>>>(A.d, A.nx, A.ny, B.d, B.nx, B.ny, d_tab + off[0],
                                                                       d_tab + off[1], d_tab + off[2], d_tab + off[3],
                                                                       d_tab + off[4], d_tab + off[5], nx, ny, nz, zchunk,
                                                                       zlo, zhi, thr_f, partials, ticket, dout, ba);
  prof_end(MODE == 0 ? PK_CC : PK_BINARY);
  VT_LAUNCHED();
  return 0;
}

// returns 0 launched, 1 not applicable, >1 error
int k_cc_staged(const DevMap &A, const DevMap &B, int nx, int ny, int nz, const std::vector<AxisEntry> &host_tab,
                const int off[6], const AxisEntry *d_tab, float thr_f, double *partials, unsigned int *ticket,
                double *dout) {
  staged::BinaryArgs ba{};
  return launch_staged<4, 0>(A, B, nx, ny, nz, host_tab, off, d_tab, thr_f, partials, ticket, dout, ba, nullptr, nullptr);
}

// two-map op through the ring: writes the planes [zlo, zhi) in which both maps are sampled; the
// caller fills the rest with the generic kernel.  mode as in cc_staged_kernel.
int k_binary_staged(int mode, const DevMap &A, const DevMap &B, const DevCoord &C, int nx, int ny, int nz,
                    const std::vector<AxisEntry> &host_tab, const int off[6], const AxisEntry *d_tab, float *outp,
                    int *zlo, int *zhi) {
  staged::BinaryArgs ba;
  ba.outp = outp; ba.gA = A; ba.gB = B; ba.gC = C;
  switch (mode) {
    case 1: return launch_staged<4, 1>(A, B, nx, ny, nz, host_tab, off, d_tab, 0.f, nullptr, nullptr, nullptr, ba, zlo, zhi);
    case 2: return launch_staged<4, 2>(A, B, nx, ny, nz, host_tab, off, d_tab, 0.f, nullptr, nullptr, nullptr, ba, zlo, zhi);
    case 3: return launch_staged<4, 3>(A, B, nx, ny, nz, host_tab, off, d_tab, 0.f, nullptr, nullptr, nullptr, ba, zlo, zhi);
    case 4: return launch_staged<4, 4>(A, B, nx, ny, nz, host_tab, off, d_tab, 0.f, nullptr, nullptr, nullptr, ba, zlo, zhi);
    case 5: return launch_staged<4, 5>(A, B, nx, ny, nz, host_tab, off, d_tab, 0.f, nullptr, nullptr, nullptr, ba, zlo, zhi);
  }
  return 1;
}

}  // namespace vt
