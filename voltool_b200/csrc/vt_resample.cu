// vt_resample.cu -- pad/crop copy, 2x downsample, 2x cubic supersample.
#include <cstdlib>
#include "vt_device.cuh"

namespace vt {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------ pad
// VolumetricData::pad copy loop (VolumetricData.C:560-574): the window [s,e) of the NEW grid is
// filled from old voxel (g - pad_minus); everything else is zero.
__global__ void __launch_bounds__(kThreads) pad_kernel(const float *__restrict__ src, int ox, int oy,
                                                       float *__restrict__ dst, PadPlan p) {
  const long rows = (long)p.ny * p.nz;
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int gy = (int)(row % p.ny), gz = (int)(row / p.ny);
    float *out = dst + row * p.nx;
    const bool live = gy >= p.sy && gy < p.ey && gz >= p.sz && gz < p.ez;
    const float *in = src + ((long)(gz - p.pzm) * oy + (gy - p.pym)) * (long)ox - p.pxm;
    for (int gx = threadIdx.x; gx < p.nx; gx += blockDim.x)
      out[gx] = (live && gx >= p.sx && gx < p.ex) ? in[gx] : 0.0f;
  }
}

int k_pad_copy(const float *src, int ox, int oy, float *dst, const PadPlan &p) {
  VT_REQUIRE_CTX();
  long rows = (long)p.ny * p.nz;
  int grid = (int)(rows < (long)ctx().sm_count * 16 ? rows : (long)ctx().sm_count * 16);
  if (grid < 1) grid = 1;
  pad_kernel<<<grid, kThreads, 0, ctx().stream>>>(src, ox, oy, dst, p);
  VT_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ downsample
// VolumetricData::downsample, VolumetricData.C:720-743: each new voxel is the double-precision
// sum of its 2x2x2 block, added in the order (x, x+1) of row y, row y+1, then the same in plane
// z+1, times the double 1/8.
__global__ void __launch_bounds__(kThreads) downsample_kernel(const float *__restrict__ src, int ox, int oy,
                                                              float *__restrict__ dst, int nx, int ny, int nz) {
  const long rows = (long)ny * nz;
  const long oxy = (long)ox * oy;
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int gy = (int)(row % ny), gz = (int)(row / ny);
    const float *b = src + 2L * gz * oxy + 2L * gy * ox;
    for (int gx = threadIdx.x; gx < nx; gx += blockDim.x) {
      const float2 r0 = __ldcs(reinterpret_cast<const float2 *>(b) + gx);
      const float2 r1 = __ldcs(reinterpret_cast<const float2 *>(b + ox) + gx);
      const float2 r2 = __ldcs(reinterpret_cast<const float2 *>(b + oxy) + gx);
      const float2 r3 = __ldcs(reinterpret_cast<const float2 *>(b + oxy + ox) + gx);
      double Z = 0.0;
      Z += r0.x; Z += r0.y; Z += r1.x; Z += r1.y; Z += r2.x; Z += r2.y; Z += r3.x; Z += r3.y;
      dst[row * nx + gx] = (float)(Z * 0.125);
    }
  }
}

// rows whose start is not 8-byte aligned (odd old x size) take the scalar path
__global__ void __launch_bounds__(kThreads) downsample_kernel_scalar(const float *__restrict__ src, int ox, int oy,
                                                                     float *__restrict__ dst, int nx, int ny, int nz) {
  const long rows = (long)ny * nz;
  const long oxy = (long)ox * oy;
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int gy = (int)(row % ny), gz = (int)(row / ny);
    const float *b = src + 2L * gz * oxy + 2L * gy * ox;
    for (int gx = threadIdx.x; gx < nx; gx += blockDim.x) {
      const float *q = b + 2 * gx;
      double Z = 0.0;
      Z += q[0]; Z += q[1]; Z += q[ox]; Z += q[ox + 1];
      Z += q[oxy]; Z += q[oxy + 1]; Z += q[oxy + ox]; Z += q[oxy + ox + 1];
      dst[row * nx + gx] = (float)(Z * 0.125);
    }
  }
}

int k_downsample(const float *src, int ox, int oy, int oz, float *dst) {
  VT_REQUIRE_CTX();
  const int nx = ox / 2, ny = oy / 2, nz = oz / 2;
  long rows = (long)ny * nz;
  if (rows < 1 || nx < 1) return 0;
  int grid = (int)(rows < (long)ctx().sm_count * 16 ? rows : (long)ctx().sm_count * 16);
  prof_begin(PK_RESAMPLE);
  if ((ox & 1) == 0)
    downsample_kernel<<<grid, kThreads, 0, ctx().stream>>>(src, ox, oy, dst, nx, ny, nz);
  else
    downsample_kernel_scalar<<<grid, kThreads, 0, ctx().stream>>>(src, ox, oy, dst, nx, ny, nz);
  prof_end(PK_RESAMPLE);
  VT_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ supersample
// VolumetricData::cubic_interp (VolumetricData.h:224-232) at mu = 1/2
__device__ __forceinline__ float cubic_half(float y0, float y1, float y2, float y3) {
  const float mu = 0.5f, mu2 = mu * mu;
  const float a0 = y3 - y2 - y0 + y1;
  const float a1 = y0 - y1 - a0;
  const float a2 = y2 - y0;
  return (a0 * mu * mu2 + a1 * mu2 + a2 * mu + y1);
}

// value at NEW x index X of an x-supersampled source row (VolumetricData.C:783-838): even X copy
// the source; odd X = 2g+1 interpolate (g-1, g, g+1, g+2) with the end sample replicated
__device__ __forceinline__ float xline(const float *__restrict__ row, int ox, int X) {
  const int g = X >> 1;
  if ((X & 1) == 0) return __ldg(row + g);
  const int gm = g > 0 ? g - 1 : 0;
  const int gp = g + 2 < ox ? g + 2 : ox - 1;
  return cubic_half(__ldg(row + gm), __ldg(row + g), __ldg(row + g + 1), __ldg(row + gp));
}

// pass 1: every EVEN output plane (Z = 2gz): x then y interpolation (:783-875).  The y stencil is
// evaluated on x-interpolated rows, exactly as the reference's in-place sweeps leave them.  A CTA
// owns kYB source rows of one plane: it stages the x-interpolated lines of those rows (+1 below,
// +2 above, clamped) in shared memory once, then every output row is either a copy (even Y) or one
// cubic over four staged lines (odd Y).
constexpr int kYB = 8;

__global__ void __launch_bounds__(kThreads) supersample_xy_kernel(const float *__restrict__ src, int ox, int oy,
                                                                  int oz, float *__restrict__ dst, int nx, int ny) {
  extern __shared__ float lines[];  // (kYB + 3) x nx
  const int yblocks = (oy + kYB - 1) / kYB;
  const long work = (long)yblocks * oz;
  const long oxy = (long)ox * oy, nxy = (long)nx * ny;
  for (long w = blockIdx.x; w < work; w += gridDim.x) {
    const int yb = (int)(w % yblocks), gz = (int)(w / yblocks);
    const int g0 = yb * kYB;
    const float *plane = src + gz * oxy;
    // staged line l <-> source row clamp(g0 - 1 + l)
    for (int l = 0; l < kYB + 3; ++l) {
      int g = g0 - 1 + l;
      g = g < 0 ? 0 : (g > oy - 1 ? oy - 1 : g);
      const float *row = plane + (long)g * ox;
      // thread t makes the pair (2t, 2t+1): no even/odd divergence, the four source values are shared
      float *ln = lines + l * nx;
      for (int t = threadIdx.x; t < ox; t += blockDim.x) {
        const float v1 = __ldg(row + t);
        ln[2 * t] = v1;
        if (t + 1 < ox) {
          const float v0 = __ldg(row + (t > 0 ? t - 1 : 0)), v2 = __ldg(row + t + 1);
          const float v3 = __ldg(row + (t + 2 < ox ? t + 2 : ox - 1));
          ln[2 * t + 1] = cubic_half(v0, v1, v2, v3);
        }
      }
    }
    __syncthreads();
    float *out = dst + 2L * gz * nxy;
    for (int r = 0; r < 2 * kYB; ++r) {
      const int Y = 2 * g0 + r;
      if (Y >= ny) break;
      const int g = Y >> 1;  // source row below / at this output row; its line is l = g - g0 + 1
      float *o = out + (long)Y * nx;
      if ((Y & 1) == 0) {
        const float *ln = lines + (g - g0 + 1) * nx;
        for (int X = threadIdx.x; X < nx; X += blockDim.x) o[X] = ln[X];
      } else {
        // rows g-1, g, g+1, g+2 with the end rows replicated (:841-875); the staged lines are already
        // clamped at the volume edges, so the neighbours are simply l-1 .. l+2
        const int l = g - g0 + 1;
        const float *l0 = lines + (l - 1) * nx, *l1 = lines + l * nx, *l2 = lines + (l + 1) * nx, *l3 = lines + (l + 2) * nx;
        for (int X = threadIdx.x; X < nx; X += blockDim.x) o[X] = cubic_half(l0[X], l1[X], l2[X], l3[X]);
      }
    }
    __syncthreads();
  }
}

// pass 2: every ODD output plane from the four surrounding even planes (:878-913).  A thread owns
// one (X,Y) column of a z chunk and slides a 4-plane window along z, so each even plane is read
// once per chunk.
constexpr int kZB = 64;

__global__ void __launch_bounds__(kThreads) supersample_z_kernel(float *__restrict__ dst, int nx, int ny, int oz) {
  const long nxy = (long)nx * ny;
  const int zchunks = (oz - 1 + kZB - 1) / kZB;
  const long work = nxy * zchunks;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < work; i += stride) {
    const int zc = (int)(i / nxy);
    const long r = i - (long)zc * nxy;
    const int gbeg = zc * kZB, gend = min(oz - 1, gbeg + kZB);
    float *col = dst + r;
    // window over source planes g-1, g, g+1, g+2 (clamped)
    float vm = col[2L * (gbeg > 0 ? gbeg - 1 : 0) * nxy];
    float v0 = col[2L * gbeg * nxy];
    float v1 = col[2L * (gbeg + 1) * nxy];
    int g = gbeg;
    for (; g + 4 <= gend; g += 4) {  // four independent loads ahead of the dependent window shifts
      float w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gp = g + q + 2 < oz ? g + q + 2 : oz - 1;
        w[q] = col[2L * gp * nxy];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        col[(2L * (g + q) + 1) * nxy] = cubic_half(vm, v0, v1, w[q]);
        vm = v0; v0 = v1; v1 = w[q];
      }
    }
    for (; g < gend; ++g) {
      const int gp = g + 2 < oz ? g + 2 : oz - 1;
      const float v2 = gp == g + 1 ? v1 : col[2L * gp * nxy];
      col[(2L * g + 1) * nxy] = cubic_half(vm, v0, v1, v2);
      vm = v0; v0 = v1; v1 = v2;
    }
  }
}

// ---- all three sweeps in one kernel --------------------------------------------------------------
// A CTA owns a tile of 64 x 8 source columns/rows (128 x 16 outputs) and marches a chunk of source
// planes.  Per source plane: (X) the x sweep of the tile's rows (+1/+2 halo rows) into shared memory,
// (Y) the y sweep from there into a ring of the last four xy-supersampled planes and out to the even
// output plane, (Z) the odd output plane between the two middle ring planes.  The intermediate even
// planes are never read back from HBM: traffic is the source once plus the output once
// (VolumetricData.C:783-913; same cubic_half, same x -> y -> z order, replicated ends by clamping).
constexpr int kSX = 64, kSY = 8, kSZ = 64, kOX = 2 * kSX, kOY = 2 * kSY;

__global__ void __launch_bounds__(kThreads) supersample_fused_kernel(const float *__restrict__ src, int ox, int oy,
                                                                     int oz, float *__restrict__ dst, int nx, int ny,
                                                                     int xt, int yt) {
  __shared__ float lines[(kSY + 3) * kOX];
  __shared__ float ring[4][kOY * kOX];
  const int tid = threadIdx.x;
  const long w = blockIdx.x;
  const int tx0 = (int)(w % xt) * kSX, g0 = (int)((w / xt) % yt) * kSY;
  const int gbeg = (int)(w / ((long)xt * yt)) * kSZ, gend = min(oz, gbeg + kSZ);
  const int X0 = 2 * tx0, Y0 = 2 * g0;
  const long oxy = (long)ox * oy, nxy = (long)nx * ny;
  for (int p = gbeg - 1; p <= gend + 1; ++p) {
    const int pc = p < 0 ? 0 : (p > oz - 1 ? oz - 1 : p);   // replicated end planes (:878-913)
    const float *plane = src + pc * oxy;
    const int slot = (p - (gbeg - 1)) & 3;
    // ---- X: staged line l <-> source row clamp(g0 - 1 + l); thread item makes the pair (2t, 2t+1)
    for (int idx = tid; idx < (kSY + 3) * kSX; idx += kThreads) {
      const int l = idx / kSX, t = idx - l * kSX;
      const int gx = tx0 + t;
      if (gx >= ox) continue;
      int g = g0 - 1 + l;
      g = g < 0 ? 0 : (g > oy - 1 ? oy - 1 : g);
      const float *row = plane + (long)g * ox;
      const float v1 = __ldg(row + gx);
      lines[l * kOX + 2 * t] = v1;
      if (gx + 1 < ox) {
        const float v0 = __ldg(row + (gx > 0 ? gx - 1 : 0)), v2 = __ldg(row + gx + 1);
        const float v3 = __ldg(row + (gx + 2 < ox ? gx + 2 : ox - 1));
        lines[l * kOX + 2 * t + 1] = cubic_half(v0, v1, v2, v3);
      }
    }
    __syncthreads();
    // ---- Y: output rows Y0 .. Y0+15 of this plane -> ring (and the even output plane 2p)
    const bool own = p >= gbeg && p < gend;
    float *even = dst + 2L * pc * nxy;
    for (int idx = tid; idx < kOY * kOX; idx += kThreads) {
      const int r = idx / kOX, X = idx - r * kOX;
      const int l = (r >> 1) + 1;
      const float *ln = lines + l * kOX + X;
      const float val = (r & 1) == 0 ? ln[0] : cubic_half(ln[-kOX], ln[0], ln[kOX], ln[2 * kOX]);
      ring[slot][idx] = val;
      if (own && Y0 + r < ny && X0 + X < nx) even[(long)(Y0 + r) * nx + X0 + X] = val;
    }
    __syncthreads();
    // ---- Z: odd plane 2g+1 from ring planes g-1, g, g+1, g+2 with g+2 = p
    const int g = p - 2;
    if (g >= gbeg && g < gend && g <= oz - 2) {
      float *odd = dst + (2L * g + 1) * nxy;
      const float *r0 = ring[(slot + 1) & 3], *r1 = ring[(slot + 2) & 3], *r2 = ring[(slot + 3) & 3], *r3 = ring[slot];
      for (int idx = tid; idx < kOY * kOX; idx += kThreads) {
        const int r = idx / kOX, X = idx - r * kOX;
        if (Y0 + r < ny && X0 + X < nx) odd[(long)(Y0 + r) * nx + X0 + X] = cubic_half(r0[idx], r1[idx], r2[idx], r3[idx]);
      }
    }
    // the next X stage only writes `lines`; its barrier orders this Z stage before the next ring write
  }
}

// ---- all three sweeps in one kernel, z window in registers, packed arithmetic -------------------------
// Same tile as above (64 x 8 source columns/rows -> 128 x 16 outputs, a chunk of source planes), but a
// thread owns four output rows of two output columns (X, X + 64) for the whole march: per source plane
// it reads the five staged x-lines it needs, forms its xy-supersampled values (2 copies, 2 cubics per
// column), keeps the last four planes of them in registers (rotating names, no moves) and writes the
// even plane and the odd plane between the two middle window planes.  The x-lines are double
// buffered: one barrier per plane, no shared-memory ring.  Every cubic runs on a packed pair
// (FADD2 / FMUL2 / FFMA2, sm_100): the two columns of a thread, or two stencils of the x sweep.
//
// cubic_half in packed form: the reference's a0*mu*mu2 + a1*mu2 + a2*mu + y1 (VolumetricData.h:224-232)
// multiplies by exact powers of two, so (a0*.5)*.25 = a0*.125 and the fused multiply-adds round exactly
// like the separate multiply and add -- unless a product is subnormal.  That cannot happen when every
// nonzero |voxel| >= 2^-100 (all differences are then multiples of 2^-123); the x sweep checks this on
// every source voxel it loads and raises `tiny_flag` otherwise, and the caller redoes the map with the
// scalar kernel.  The order of the sweeps is the reference's (x, then y on x-swept rows, then z).
__device__ __forceinline__ f32x2 cubic_half2(f32x2 y0, f32x2 y1, f32x2 y2, f32x2 y3, f32x2 c125, f32x2 c25, f32x2 c5) {
  const f32x2 a0 = add2(sub2(sub2(y3, y2), y0), y1);
  const f32x2 a1 = sub2(sub2(y0, y1), a0);
  const f32x2 a2 = sub2(y2, y0);
  f32x2 s = fma2(a1, c25, mul2(a0, c125));
  s = fma2(a2, c5, s);
  return add2(s, y1);
}

constexpr int kMLines = kSY + 3;
constexpr int kMPairs = kMLines * kSX / 2;                       // x-sweep items are done two at a time
constexpr int kRawW = kSX + 3, kRawP = kSX + 4;                   // raw tile: columns tx0-1 .. tx0+kSX+1, row pitch
constexpr int kRawN = kMLines * kRawW;                            // elements staged per plane
constexpr int kRawStages = 4;                                     // planes in flight: kRawStages - 1
// RPT output rows per thread: 4 (256 threads, 3 CTAs per SM) or 2 (512 threads, 2 CTAs per SM: a third more warps)
constexpr int kMarchRows = 2;

__device__ __forceinline__ void cp_async4(float *smem_dst, const void *gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int RPT>
__global__ void __launch_bounds__(kSX * (kOY / RPT), RPT == 4 ? 3 : 2)
    supersample_march_kernel(const float *__restrict__ src, int ox, int oy, int oz, float *__restrict__ dst, int nx, int ny, int xt,
                             int yt, int *__restrict__ tiny_flag) {
  constexpr int kThreads = kSX * (kOY / RPT);                        // (shadows the file-wide 256)
  constexpr int kMIt = (kMPairs + kThreads - 1) / kThreads;
  constexpr int kRawIt = (kRawN + kThreads - 1) / kThreads;
  constexpr int NL = RPT / 2 + 3;                                    // x-lines a thread reads per plane
  __shared__ __align__(16) float lines[2][kMLines * kOX];
  __shared__ float raw[kRawStages][kMLines * kRawP];
  const int tid = threadIdx.x;
  const long w = blockIdx.x;
  const int tx0 = (int)(w % xt) * kSX, g0 = (int)((w / xt) % yt) * kSY;
  const int gbeg = (int)(w / ((long)xt * yt)) * kSZ, gend = min(oz, gbeg + kSZ);
  const int X0 = 2 * tx0, Y0 = 2 * g0;
  const long oxy = (long)ox * oy, nxy = (long)nx * ny;
  const int nsteps = gend - gbeg + 3;   // source planes gbeg-1 .. gend+1 (clamped: ends replicated, :878-913)
  const f32x2 c125 = pack2(0.125f, 0.125f), c25 = pack2(0.25f, 0.25f), c5 = pack2(0.5f, 0.5f);

  // ---- raw-tile staging: every thread copies up to kRawIt source voxels per plane with cp.async, rows and
  // columns clamped to the map (so the stencils of the x sweep need no edge cases), planes ahead of their use
  unsigned long long rp[kRawIt];   // source address of the element in the plane staged next
  int rdst[kRawIt];                // shared-memory slot inside a stage, < 0: none
#pragma unroll
  for (int k = 0; k < kRawIt; ++k) {
    const int e = tid + kThreads * k;
    const int l = e / kRawW, c = e - l * kRawW;
    int g = g0 - 1 + l, gx = tx0 - 1 + c;
    g = g < 0 ? 0 : (g > oy - 1 ? oy - 1 : g);
    gx = gx < 0 ? 0 : (gx > ox - 1 ? ox - 1 : gx);
    rdst[k] = e < kRawN ? l * kRawP + c : -1;
    const int pfirst = gbeg > 0 ? gbeg - 1 : 0;
    rp[k] = (unsigned long long)(src + (pfirst * oxy + (long)g * ox + gx));
  }
  const unsigned long long sstep = (unsigned long long)oxy * sizeof(float);
  int staged = 0;   // planes issued so far; the next one is p = gbeg - 1 + staged
  auto stage_plane = [&]() {
    if (staged < nsteps) {
      float *buf = raw[staged & (kRawStages - 1)];
#pragma unroll
      for (int k = 0; k < kRawIt; ++k)
        if (rdst[k] >= 0) cp_async4(buf + rdst[k], reinterpret_cast<const void *>(rp[k]));
      const int p = gbeg - 1 + staged;
      if (p >= 0 && p < oz - 1) {
#pragma unroll
        for (int k = 0; k < kRawIt; ++k) rp[k] += sstep;
      }
    }
    cp_async_commit();
    ++staged;
  };

  // ---- x sweep items (two per trip): raw-tile offset of the left stencil sample and the shared-memory slot
  // of the output pair (2t, 2t+1); slot < 0: the item lies right of the map
  int roff[kMIt][2], sdst[kMIt][2];
#pragma unroll
  for (int it = 0; it < kMIt; ++it)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = tid + kThreads * it;
      const int idx = (q < kMPairs ? q : 0) + h * kMPairs;
      const int l = idx / kSX, t = idx - l * kSX;
      roff[it][h] = l * kRawP + t;
      sdst[it][h] = (q < kMPairs && tx0 + t < ox) ? l * kOX + 2 * t : -1;
    }
  unsigned int tiny = 0;
  auto x_sweep = [&](int s) {
    const float *rw = raw[s & (kRawStages - 1)];
    float *ln = lines[s & 1];
#pragma unroll
    for (int it = 0; it < kMIt; ++it) {
      if (kThreads * it + tid >= kMPairs) continue;
      const float *ra = rw + roff[it][0], *rb = rw + roff[it][1];
      const float a1 = ra[1], b1 = rb[1];
      const f32x2 o = cubic_half2(pack2(ra[0], rb[0]), pack2(a1, b1), pack2(ra[2], rb[2]), pack2(ra[3], rb[3]), c125, c25, c5);
      float oa, ob;
      unpack2(o, oa, ob);
      tiny |= (unsigned int)(((__float_as_uint(a1) & 0x7fffffffu) - 1u) < 0x0d7fffffu);
      tiny |= (unsigned int)(((__float_as_uint(b1) & 0x7fffffffu) - 1u) < 0x0d7fffffu);
      if (sdst[it][0] >= 0) *reinterpret_cast<float2 *>(ln + sdst[it][0]) = make_float2(a1, oa);
      if (sdst[it][1] >= 0) *reinterpret_cast<float2 *>(ln + sdst[it][1]) = make_float2(b1, ob);
    }
  };

  // ---- y / z sweeps: output columns X and X + 64, RPT output rows RPT rg ..
  const int X = tid & (kSX - 1), rg = tid >> 6;
  unsigned int smask = 0;   // bit j: row j of column X is inside the map; bit RPT + j: column X + 64
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const bool rowok = Y0 + RPT * rg + j < ny;
    if (rowok && X0 + X < nx) smask |= 1u << j;
    if (rowok && X0 + X + kSX < nx) smask |= (1u << RPT) << j;
  }
  // this thread's four output rows in the even plane 2p of the current step (only dereferenced when
  // that plane exists); the odd plane written in the same step, 2(p-2)+1, lies 3 planes below
  unsigned long long oe[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) oe[j] = (unsigned long long)(dst + (2L * (gbeg - 1) * nxy + (long)(Y0 + RPT * rg + j) * nx + X0 + X));
  const unsigned long long ostep = 2ull * nxy * sizeof(float), oback = 3ull * nxy * sizeof(float);
  f32x2 A[RPT], B[RPT], C[RPT], D[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) A[j] = B[j] = C[j] = D[j] = 0ull;

  auto plane_step = [&](int s, const f32x2 (&w0)[RPT], const f32x2 (&w1)[RPT], const f32x2 (&w2)[RPT], f32x2 (&w3)[RPT]) {
    const int p = gbeg - 1 + s;
    cp_async_wait<kRawStages - 3>();   // this thread's copies of plane s+1 have landed ...
    __syncthreads();      // ... and everybody's; lines[s&1] is complete; lines[(s+1)&1] and the oldest raw stage are free
    if (s + 1 < nsteps) x_sweep(s + 1);
    const float *ln = lines[s & 1];
    f32x2 L[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k)
      L[k] = pack2(ln[((RPT / 2) * rg + k) * kOX + X], ln[((RPT / 2) * rg + k) * kOX + X + kSX]);
    stage_plane();        // plane s + kRawStages - 1
#pragma unroll
    for (int i = 0; i < RPT / 2; ++i) {
      w3[2 * i] = L[i + 1];                                                         // even output rows: copies (:783-794)
      w3[2 * i + 1] = cubic_half2(L[i], L[i + 1], L[i + 2], L[i + 3], c125, c25, c5);  // odd rows (:841-875)
    }
    if (p >= gbeg && p < gend) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        float a, b;
        unpack2(w3[j], a, b);
        float *o = reinterpret_cast<float *>(oe[j]);
        if ((smask >> j) & 1u) __stcs(o, a);
        if ((smask >> (RPT + j)) & 1u) __stcs(o + kSX, b);
      }
    }
    const int g = p - 2;   // odd plane 2g+1 from planes g-1, g, g+1, g+2 = p
    if (g >= gbeg && g < gend && g <= oz - 2) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        float a, b;
        unpack2(cubic_half2(w0[j], w1[j], w2[j], w3[j], c125, c25, c5), a, b);
        float *o = reinterpret_cast<float *>(oe[j] - oback);
        if ((smask >> j) & 1u) __stcs(o, a);
        if ((smask >> (RPT + j)) & 1u) __stcs(o + kSX, b);
      }
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) oe[j] += ostep;
  };
#pragma unroll
  for (int k = 0; k < kRawStages - 1; ++k) stage_plane();
  cp_async_wait<kRawStages - 2>();
  __syncthreads();
  x_sweep(0);
  int s = 0;
  while (true) {
    plane_step(s, A, B, C, D); if (++s >= nsteps) break;
    plane_step(s, B, C, D, A); if (++s >= nsteps) break;
    plane_step(s, C, D, A, B); if (++s >= nsteps) break;
    plane_step(s, D, A, B, C); if (++s >= nsteps) break;
  }
  if (tiny) atomicOr(tiny_flag, 1);
}

int k_supersample(const float *src, int ox, int oy, int oz, float *dst) {
  VT_REQUIRE_CTX();
  if (ox < 3 || oy < 3 || oz < 3) { set_error("vt_supersample: every dimension must be >= 3"); return 1; }
  const int nx = 2 * ox - 1, ny = 2 * oy - 1;
  long rows = (long)ny * oz;
  int grid = (int)(rows < (long)ctx().sm_count * 16 ? rows : (long)ctx().sm_count * 16);
  prof_begin(PK_RESAMPLE);
  // measured on B200: the single kernel wins once the intermediate even planes no longer fit L2
  // (512^3 -> 1023^3: 2.06 vs 2.51 ms); below that the two-pass form is as fast or faster
  const char *mode = getenv("VOLTOOL_SUPERSAMPLE_FUSED");
  const int fused = mode ? atoi(mode) : ((long)ox * oy * oz >= (48L << 20) ? 2 : 0);
  if (fused == 2 && (long)ox * oy <= 0x7fffffffL - 4) {
    const int xt = (ox + kSX - 1) / kSX, yt = (oy + kSY - 1) / kSY;
    const long work = (long)xt * yt * ((oz + kSZ - 1) / kSZ);
    if (work <= 0x7fffffffL) {
      // the packed kernel is exact unless a nonzero voxel is tinier than 2^-100: checked on the device,
      // such a map is redone below by the scalar kernels
      static int *d_flag = nullptr, *h_flag = nullptr;
      if (!d_flag) {
        VT_CUDA(cudaMalloc((void **)&d_flag, sizeof(int)));
        VT_CUDA(cudaMallocHost((void **)&h_flag, sizeof(int)));
      }
      VT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), ctx().stream));
      supersample_march_kernel<kMarchRows><<<(unsigned int)work, kSX * (kOY / kMarchRows), 0, ctx().stream>>>(src, ox, oy, oz, dst, nx, ny, xt, yt, d_flag);
      VT_LAUNCHED();
      VT_CUDA(cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
      VT_CUDA(cudaStreamSynchronize(ctx().stream));
      if (*h_flag == 0) {
        prof_end(PK_RESAMPLE);
        return 0;
      }
    }
  }
  if (fused != 0) {   // also the exact fallback of the packed kernel
    const int xt = (ox + kSX - 1) / kSX, yt = (oy + kSY - 1) / kSY;
    const long work = (long)xt * yt * ((oz + kSZ - 1) / kSZ);
    if (work <= 0x7fffffffL) {
      supersample_fused_kernel<<<(unsigned int)work, kThreads, 0, ctx().stream>>>(src, ox, oy, oz, dst, nx, ny, xt, yt);
      prof_end(PK_RESAMPLE);
      VT_LAUNCHED();
      return 0;
    }
  }
  const size_t smem = sizeof(float) * (size_t)(kYB + 3) * nx;
  if (smem > 200 * 1024) { set_error("vt_supersample: rows too long (%d)", ox); return 1; }
  if (smem > 48 * 1024)
    VT_CUDA(cudaFuncSetAttribute(supersample_xy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    const long work = (long)((oy + kYB - 1) / kYB) * oz;
    grid = (int)(work < (long)ctx().sm_count * 8 ? work : (long)ctx().sm_count * 8);
  }
  supersample_xy_kernel<<<grid, kThreads, smem, ctx().stream>>>(src, ox, oy, oz, dst, nx, ny);
  VT_LAUNCHED();
  supersample_z_kernel<<<ctx().sm_count * 8, kThreads, 0, ctx().stream>>>(dst, nx, ny, oz);
  prof_end(PK_RESAMPLE);
  VT_LAUNCHED();
  return 0;
}

}  // namespace vt
