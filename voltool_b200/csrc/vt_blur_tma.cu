// vt_blur_tma.cu -- VolumetricData::gaussian_blur (VolumetricData.C:977-997) in ONE pass over HBM.
//
// A CTA owns a TXW x TYH column of the map and marches along z.  Per source plane:
//   (load)  one lane of a producer warp issues a 3-D tensor-map TMA box load (cp.async.bulk.tensor.3d, completion on
//           an mbarrier) of the tile + its x/y halo, S planes ahead.  TMA zero-fills what lies outside the map; the
//           tiles on the map border overwrite those cells with the clamp-to-edge value before the x pass.
//   (x)     a thread per staged row: 16 + 2 RA floats by 16-byte shared loads, 16 outputs as 8 packed pairs.  A pair
//           (j, j+1) receives input i as tap k of j and tap k-1 of j+1: FFMA2 with the input broadcast and the weight
//           pair (w[k], w[k-1]) -- no register moves, because the pair is two OUTPUTS, not two inputs; the first tap
//           of j and the last tap of j+1 have no partner and are scalar FFMAs.  Results go to a TRANSPOSED tile
//           T[x][y] so that ...
//   (y)     ... a thread reads the 4 + 2R values of its column's window with 16-byte loads and forms the y pass of
//           its 4 rows the same way (2 pairs), and
//   (z)     scatters the two pairs into 2R+1 pending output planes held in registers (input plane zin is tap k of
//           output zin + R - k), storing the plane whose last tap just arrived.
// Per output the taps are applied in the order k = 0 .. 2R on an accumulator that starts at 0, every step one fmaf:
// exactly the chain of the oracle / of the other blur kernels, so all forms agree bit for bit.
// HBM traffic: the source once (halo re-reads are L2 hits) + the output once = 8 B/voxel.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "vt_device.cuh"
#include "vt_tma.h"

namespace vt {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int tma_encode_map3d(CUtensorMap *map, const float *base, int nx, int ny, int nz, int bx, int by, int bz, bool swizzle128) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return 1;
  const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
  const cuuint64_t strides[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
  const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

namespace tmab {

constexpr int kTaps = 2 * 8 + 1;
// weights as the kernel consumes them: c_w[k] = w[k]; c_w2[k] = (w[k], w[k-1]), the two taps one input applies to an
// (even, odd) output pair -- stored as aligned pairs so that they load straight into uniform-register pairs
__constant__ float c_w[kTaps + 1];
__constant__ float2 c_w2[kTaps + 1];

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned int bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned int dst, const CUtensorMap *map, int cx, int cy, int cz, unsigned int bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(cx), "r"(cy), "r"(cz), "r"(bar)
      : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned int addr) {
  float4 f;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(addr));
  return f;
}
__device__ __forceinline__ void sts32(unsigned int addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

template <int R, int TXW, int TYH>
struct Cfg {
  static constexpr int NT = 2 * R + 1;
  // x halo staged 8 wide on either side whatever R is: the box origin x0 - 8 keeps every TMA row start 16-byte
  // aligned in global memory (a box whose rows start at 8-byte offsets faults), and the staged row is exactly
  // 32 floats = 128 bytes = the span of SWIZZLE_128B, which makes the x pass's 16-byte reads of 8 consecutive rows
  // hit 8 different bank groups although the rows are 128 bytes apart
  static constexpr int RA = 8;
  static constexpr int OFFX = RA - R;              // window position of tap 0 of output 0
  static constexpr int BX = TXW + 2 * RA;          // staged floats per row
  static constexpr int ROWS = TYH + 2 * R;         // staged rows per plane
  static constexpr int PY0 = (ROWS + 3) & ~3;
  static constexpr int PY = ((PY0 / 4) & 1) ? PY0 : PY0 + 4;   // pitch of the transposed tile: odd number of 16-byte units
  static constexpr int STAGE_BYTES = (ROWS * BX * 4 + 1023) & ~1023;   // swizzled TMA destinations: 1024-byte aligned
  static constexpr int S = 6;                      // raw planes in flight (TMA ring)
  static constexpr int D = 9;                      // x-blurred planes: three groups of three (written one group ahead of the
                                                   // readers, into slots whose readers finished a whole group ago)
  static constexpr int NTHR = (TXW * TYH) / 4;     // a thread owns 1 column x 4 rows
  static constexpr int YWIN = 4 + 2 * R;
  static constexpr int YQ = (YWIN + 3) / 4;
  static constexpr int Q0 = OFFX / 4, Q1 = (OFFX + 16 + 2 * R - 1) / 4;   // 16-byte chunks of a staged row the x pass reads
  static constexpr int T_BYTES = TXW * PY * 4;
  static constexpr size_t SMEM = (size_t)S * STAGE_BYTES + (size_t)D * T_BYTES + 16 * (S + D) + 1024;   // (barriers: 16 bytes per group of 3)
  static_assert(TXW == 16 && TYH % 4 == 0 && R >= 1 && R <= 8 && 3 * ROWS <= NTHR, "tile shape");
  static_assert(BX * 4 == 128, "a staged row is one 128-byte swizzle span");
};

// float index of cell (row r, column c) of a staged plane under CU_TENSOR_MAP_SWIZZLE_128B: the 16-byte chunk index
// is XORed with the row index modulo 8
__device__ __forceinline__ int sw_index(int r, int c) { return r * 32 + ((((c >> 2) ^ r) & 7) << 2) + (c & 3); }

// NOUT outputs (NOUT / 2 pairs) of one line from a window of inputs held in registers.  Pair (j, j+1) receives input
// i0 + k as tap k of j and tap k-1 of j+1: one FFMA2 with the input broadcast and the weight pair (w[k], w[k-1]);
// tap 0 of j and tap 2R of j+1 have no partner and are scalar FFMAs.  Per output: taps 0 .. 2R in order, from 0.
template <int R, int NOUT, int OFF, int NIN>
__device__ __forceinline__ void conv_pairs(const float (&in)[NIN], f32x2 (&acc)[NOUT / 2]) {
#pragma unroll
  for (int p = 0; p < NOUT / 2; ++p) {
    const int i0 = OFF + 2 * p;
    float lo = __fmaf_rn(c_w[0], in[i0], 0.0f), hi = 0.0f;        // tap 0 of the even output; the odd one has not started
    f32x2 a = pack2(lo, hi);
#pragma unroll
    for (int k = 1; k <= 2 * R; ++k) a = fma2(pack2(in[i0 + k], in[i0 + k]), pack2(c_w2[k].x, c_w2[k].y), a);   // taps (k, k-1)
    unpack2(a, lo, hi);
    hi = __fmaf_rn(c_w[2 * R], in[i0 + 2 * R + 1], hi);            // tap 2R of the odd output
    acc[p] = pack2(lo, hi);
  }
}

template <int R, int S_>
__device__ __forceinline__ void z_scatter(f32x2 (&acc)[2][2 * R + 1], const f32x2 (&v)[2], bool emit, float *__restrict__ o, int nx,
                                          unsigned int okmask) {
  constexpr int NT = 2 * R + 1;
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    const int slot = (S_ + R - k + 2 * NT) % NT;   // slot of output zin + R - k, with slot(z) = z mod NT and S_ = zin mod NT
    const f32x2 wk = pack2(c_w[k], c_w[k]);
#pragma unroll
    for (int c = 0; c < 2; ++c) acc[c][slot] = fma2(wk, v[c], k == 0 ? pack2(0.0f, 0.0f) : acc[c][slot]);
  }
  if (emit) {   // output zin - R: its last tap (k = 2R) was just added
    constexpr int slot = (S_ - R + 2 * NT) % NT;
    float a, b, c, d;
    unpack2(acc[0][slot], a, b);
    unpack2(acc[1][slot], c, d);
    if (okmask == 15u) {   // interior: four unconditional stores
      __stcs(o, a); __stcs(o + nx, b); __stcs(o + 2 * nx, c); __stcs(o + 3 * nx, d);
    } else {
      if (okmask & 1u) __stcs(o, a);
      if (okmask & 2u) __stcs(o + nx, b);
      if (okmask & 4u) __stcs(o + 2 * nx, c);
      if (okmask & 8u) __stcs(o + 3 * nx, d);
    }
  }
}

template <int R>
__device__ __forceinline__ void z_dispatch(int s, f32x2 (&acc)[2][2 * R + 1], const f32x2 (&v)[2], bool emit, float *__restrict__ o,
                                           int nx, unsigned int okmask) {
  constexpr int NT = 2 * R + 1;
#define VT_ZC(S_) case S_: if constexpr (S_ < NT) z_scatter<R, (S_ < NT ? S_ : 0)>(acc, v, emit, o, nx, okmask); break;
  switch (s) {
    VT_ZC(0) VT_ZC(1) VT_ZC(2) VT_ZC(3) VT_ZC(4) VT_ZC(5) VT_ZC(6) VT_ZC(7) VT_ZC(8)
    VT_ZC(9) VT_ZC(10) VT_ZC(11) VT_ZC(12) VT_ZC(13) VT_ZC(14) VT_ZC(15) VT_ZC(16)
    default: break;
  }
#undef VT_ZC
}

__device__ __forceinline__ bool mbar_test(unsigned int bar, unsigned int parity) {
  unsigned int done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned int bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// One CTA = one (tile, z chunk): it marches its whole z range.  CTAs are numbered tile-major inside a chunk, so the
// CTAs resident at any time are neighbouring tiles at about the same z, and the halo re-reads of a plane (each source
// column is staged by two tiles in x, the y halo rows by two in y) hit L2 while that plane is still resident.
//
// No block barrier in the steady state.  Planes flow through two shared-memory rings guarded by mbarriers:
//   raw[S]  TMA boxes, issued by a ninth (producer) warp S planes ahead
//           (full: the TMA's byte count; empty: the warps that x-passed the plane)
//   T[D]    x-blurred planes, transposed (full: the warps of the x pass; empty: every warp after its y window)
// The x pass of THREE planes is spread over the whole CTA (thread t: row t % ROWS of plane 3g + t / ROWS), one
// group of three planes ahead of the y / z work, so every thread carries the same load and warps only ever wait
// for data, not for each other.  What this costs and why it is the default only for short kernels: DESIGN.md 3.5.
// warps that hold x-role rows of the planes q with q % 3 == k (threads k ROWS .. (k + 1) ROWS - 1)
template <int ROWS>
__host__ __device__ constexpr int x_warps(int k) { return ((k + 1) * ROWS - 1) / 32 - (k * ROWS) / 32 + 1; }

template <int R, int TXW, int TYH, bool USE_TMA>
__global__ void __launch_bounds__(Cfg<R, TXW, TYH>::NTHR + 32, 2)
    blur_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ src, float *__restrict__ out, int nx, int ny,
                    int nz, int xt, int tiles, int zchunk, int abl) {
  using C = Cfg<R, TXW, TYH>;
  constexpr int NT = C::NT, BX = C::BX, ROWS = C::ROWS, PY = C::PY, S = C::S, D = C::D, NTHR = C::NTHR;
  static_assert(S % 3 == 0 && D % 3 == 0, "a ring slot always holds planes of one x-role group");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const unsigned int sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;          // shared-window byte addresses throughout
  const unsigned int tbase = sbase + S * C::STAGE_BYTES;
  const unsigned int bars = tbase + D * C::T_BYTES;
  // every barrier counts WARPS, not threads (an arrive per thread made the handshakes alone cost 0.26 ms at 512^3): the
  // lanes of a warp that share a plane meet at a __syncwarp and one of them arrives
  const unsigned int raw_full = bars, raw_empty = bars + 8 * S, t_full = bars + 16 * S, t_empty = bars + 16 * S + 8 * D;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(raw_full + 8 * s, 1); mbar_init(raw_empty + 8 * s, x_warps<ROWS>(s % 3)); }
    for (int s = 0; s < D; ++s) { mbar_init(t_full + 8 * s, x_warps<ROWS>(s % 3)); mbar_init(t_empty + 8 * s, NTHR / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int item = blockIdx.x;
  const int t = item % tiles, zc = item / tiles;
  const int tx = t % xt, ty = t / xt;
  const int x0 = tx * TXW, y0 = ty * TYH;
  const int zlo = zc * zchunk, zhi = min(nz, zlo + zchunk);
  const int pf = max(zlo - R, 0), pl = min(zhi - 1 + R, nz - 1);
  const int nplanes = pl - pf + 1;
  // staged columns outside the map (TMA wrote zeros there): [0, cx0) and [cx1, BX); cx0 is 0 or 8, cx1 a multiple of 4
  const int cx0 = max(0, C::RA - x0), cx1 = min(BX, nx - (x0 - C::RA));
  const bool xborder = cx0 > 0 || cx1 < BX;
  // x role: staged row xrow of the plane 3 g + xq of every group g of three planes
  const int xq = tid / ROWS, xrow = tid - xq * ROWS;
  const bool xrole = tid < 3 * ROWS;
  const int xsrc = min(max(y0 - R + xrow, 0), ny - 1) - (y0 - R);   // clamp to edge in y: read the staged row of the clamped y
  // y / z role: column lx of the tile, rows 4 rg .. 4 rg + 3
  const int lx = tid % TXW, rg = tid / TXW;
  const int gx = x0 + lx, gy = y0 + 4 * rg;
  unsigned int okmask = 0;
  if (gx < nx) {
#pragma unroll
    for (int j = 0; j < 4; ++j) okmask |= (gy + j < ny) ? (1u << j) : 0u;
  }
  float *o = out + ((size_t)zlo * ny + gy) * nx + gx;
  const size_t nxy = (size_t)nx * ny;
  const unsigned int ywin = tbase + (unsigned int)(lx * PY + 4 * rg) * 4u;   // this thread's y window in T slot 0

  if (tid >= NTHR) {   // producer warp: one lane keeps the TMA ring full, S planes ahead of the x pass
    if (USE_TMA && !(abl & 16) && tid == NTHR) {
      for (int i = 0; i < nplanes; ++i) {
        const unsigned int st = (unsigned int)i % S, par = (((unsigned int)i / S) & 1u) ^ 1u;
        mbar_wait(raw_empty + 8 * st, par);             // every warp that x-passed the stage's last plane has arrived
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(raw_full + 8 * st, (unsigned int)(BX * ROWS * sizeof(float)));
        tma_load_3d(sbase + st * C::STAGE_BYTES, &tmap, x0 - C::RA, y0 - R, pf + i, raw_full + 8 * st);
      }
    }
    return;
  }
  // the lanes of this warp that x-pass rows of the same plane, and the one that arrives for them
  const unsigned int xmask = __match_any_sync(0xffffffffu, xrole ? xq : -1);
  const bool xlead = xrole && (tid & 31) == __ffs(xmask) - 1;
  const bool wlead = (tid & 31) == 0;
  // x pass of plane q (this thread's row) into T slot q % D
  auto x_phase = [&](int q) {
    const unsigned int stg = (unsigned int)q % S, slot = (unsigned int)q % D;
    float in[BX];
    if (USE_TMA) {
      if (!(abl & 16)) mbar_wait(raw_full + 8 * stg, ((unsigned int)q / S) & 1u);
      // chunk c of staged row r lives at byte (r * 128) | ((c ^ (r & 7)) << 4): one XOR per chunk
      const unsigned int ra = (sbase + stg * C::STAGE_BYTES + (unsigned int)xsrc * 128u) | (((unsigned int)xsrc & 7u) << 4);
#pragma unroll
      for (int qd = C::Q0; qd <= C::Q1; ++qd) {
        const float4 f = lds128(ra ^ ((unsigned int)qd << 4));
        in[4 * qd] = f.x; in[4 * qd + 1] = f.y; in[4 * qd + 2] = f.z; in[4 * qd + 3] = f.w;
      }
      __syncwarp(xmask);                             // the rows are in registers: the stage may be refilled
      if (xlead && !(abl & 16)) mbar_arrive(raw_empty + 8 * stg);
      if (xborder) {                                 // clamp to edge in x, in registers
        if (cx0 > 0) {
#pragma unroll
          for (int c = 0; c < C::RA; ++c) in[c] = in[C::RA];
        }
        if (cx1 < BX) {
          float edge = in[11];
#pragma unroll
          for (int c = 15; c < BX; c += 4) edge = (c < cx1) ? in[c] : edge;   // in[cx1 - 1]
#pragma unroll
          for (int c = 12; c < BX; ++c) in[c] = (c >= cx1) ? edge : in[c];
        }
      }
    } else {   // fallback staging without TMA: the thread reads its row straight from global memory, clamped
      const float *rowp = src + ((size_t)(pf + q) * ny + (size_t)(y0 - R + xsrc)) * nx;
#pragma unroll
      for (int c = C::Q0 * 4; c < C::Q1 * 4 + 4; ++c) in[c] = __ldg(rowp + min(max(x0 - C::RA + c, 0), nx - 1));
    }
    f32x2 a[8];
    if (abl & 4) {
#pragma unroll
      for (int p = 0; p < 8; ++p) a[p] = pack2(in[2 * p + 8], in[2 * p + 9]);
    } else conv_pairs<R, 16, C::OFFX, BX>(in, a);
    mbar_wait(t_empty + 8 * slot, (((unsigned int)q / D) & 1u) ^ 1u);   // every reader of the plane D before q is done
    const unsigned int ta = tbase + slot * C::T_BYTES + (unsigned int)xrow * 4u;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      float lo, hi;
      unpack2(a[p], lo, hi);
      sts32(ta + (2 * p) * PY * 4, lo);
      sts32(ta + (2 * p + 1) * PY * 4, hi);
    }
    __syncwarp(xmask);
    if (xlead) mbar_arrive(t_full + 8 * slot);
  };

  f32x2 acc[2][NT];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int k = 0; k < NT; ++k) acc[c][k] = pack2(0.0f, 0.0f);
  f32x2 v[2] = {pack2(0.0f, 0.0f), pack2(0.0f, 0.0f)};

  if (xrole && xq < nplanes) x_phase(xq);           // group 0
  int s = ((zlo - R) % NT + NT) % NT;
  int p = 0;                                        // next plane to consume
  int p3 = 0;                                       // p % 3
  for (int zin = zlo - R; zin < zhi + R; ++zin) {
    const int zp = zin < 0 ? 0 : (zin > nz - 1 ? nz - 1 : zin);   // clamp to edge: virtual planes repeat the border plane
    if (zp - pf == p) {                             // a new plane: uniform over the CTA
      if (p3 == 0 && xrole && p + 3 + xq < nplanes) x_phase(p + 3 + xq);   // the next group of three planes
      const unsigned int slot = (unsigned int)p % D;
      mbar_wait(t_full + 8 * slot, ((unsigned int)p / D) & 1u);
      float win[4 * C::YQ];
#pragma unroll
      for (int qd = 0; qd < C::YQ; ++qd) {
        const float4 f = lds128(ywin + slot * C::T_BYTES + 16u * qd);
        win[4 * qd] = f.x; win[4 * qd + 1] = f.y; win[4 * qd + 2] = f.z; win[4 * qd + 3] = f.w;
      }
      __syncwarp();
      if (wlead) mbar_arrive(t_empty + 8 * slot);
      if (abl & 8) { v[0] = pack2(win[0], win[1]); v[1] = pack2(win[2], win[3]); }
      else conv_pairs<R, 4, 0, 4 * C::YQ>(win, v);
      ++p;
      p3 = p3 == 2 ? 0 : p3 + 1;
    }
    const bool emit = zin - R >= zlo;               // the output completed by this step
    if (abl & 2) {
      if (emit && !(abl & 1) && okmask == 15u) { float a_, b_; unpack2(v[0], a_, b_); __stcs(o, a_); __stcs(o + nx, b_); unpack2(v[1], a_, b_); __stcs(o + 2 * nx, a_); __stcs(o + 3 * nx, b_); }
    } else z_dispatch<R>(s, acc, v, emit && !(abl & 1), o, nx, okmask);
    if (emit) o += nxy;
    s = s + 1 == NT ? 0 : s + 1;
  }
}

template <int R, int TXW, int TYH>
static int launch(const float *src, float *dst, int nx, int ny, int nz) {
  using C = Cfg<R, TXW, TYH>;
  Ctx &c = ctx();
  CUtensorMap map;
  if (tma_encode_map3d(&map, src, nx, ny, nz, C::BX, C::ROWS, 1, true)) return 1;
  const int xt = (nx + TXW - 1) / TXW, yt = (ny + TYH - 1) / TYH;
  const long tiles = (long)xt * yt;
  static int ctas_per_sm = 0;
  const bool no_tma = getenv("VOLTOOL_BLUR_NOTMA") != nullptr;   // debugging: stage with plain loads
  auto kern = no_tma ? blur_tma_kernel<R, TXW, TYH, false> : blur_tma_kernel<R, TXW, TYH, true>;
  if (!ctas_per_sm) {
    VT_CUDA(cudaFuncSetAttribute(blur_tma_kernel<R, TXW, TYH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    VT_CUDA(cudaFuncSetAttribute(blur_tma_kernel<R, TXW, TYH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    VT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, C::NTHR + 32, C::SMEM));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  // z chunks: one CTA per (tile, chunk).  More chunks fill the machine when there are few tiles, and even out the
  // last wave when there are many; every chunk pays 2R warm-up planes.  Pick the count with the best product of
  // wave efficiency and useful-plane fraction.
  const long slots = (long)c.sm_count * ctas_per_sm;
  int best_zc = 1;
  double best_eff = 0.0;
  for (int zc = 1; zc <= 64; ++zc) {
    const int len = (nz + zc - 1) / zc;
    if (zc > 1 && len < 4 * R + 4) break;
    const long items = tiles * ((nz + len - 1) / len);
    const long waves = (items + slots - 1) / slots;
    const double eff = (double)items / (double)(waves * slots) * (double)len / (double)(len + 2 * R);
    if (eff > best_eff * 1.02) { best_eff = eff; best_zc = zc; }
  }
  if (const char *e = getenv("VOLTOOL_BLUR_ZCHUNKS")) best_zc = std::max(1, atoi(e));
  const int zchunk = (nz + best_zc - 1) / best_zc;
  const long items = tiles * ((nz + zchunk - 1) / zchunk);
  if (items > 0x7fffffffL) return 1;
  kern<<<(unsigned int)items, C::NTHR + 32, C::SMEM, c.stream>>>(map, src, dst, nx, ny, nz, xt, (int)tiles, zchunk,
                                                            getenv("VOLTOOL_BLUR_ABL") ? atoi(getenv("VOLTOOL_BLUR_ABL")) : 0);
  VT_LAUNCHED();
  return 0;
}

}  // namespace tmab

// returns 0 launched, 1 not applicable (caller falls back), >1 error
int k_blur_tma(const float *src, float *dst, int nx, int ny, int nz, const float *w, int R) {
  using namespace tmab;
  if (R < 1 || R > 8 || (nx & 3) != 0 || nx < 4 || ((uintptr_t)src & 15) != 0) return 1;
  if ((long)nx * ny * 4 >= (1L << 40)) return 1;
  Ctx &c = ctx();
  float hw[kTaps + 1] = {0};
  float2 hw2[kTaps + 1];
  for (int k = 0; k <= 2 * R; ++k) { hw[k] = w[k]; hw2[k] = make_float2(w[k], k ? w[k - 1] : 0.0f); }
  VT_CUDA(cudaMemcpyToSymbolAsync(c_w, hw, sizeof(float) * (2 * R + 1), 0, cudaMemcpyHostToDevice, c.stream));
  VT_CUDA(cudaMemcpyToSymbolAsync(c_w2, hw2, sizeof(float2) * (2 * R + 1), 0, cudaMemcpyHostToDevice, c.stream));
#define VT_BL(RR) case RR: return launch<RR, 16, 64>(src, dst, nx, ny, nz);
  switch (R) {
    VT_BL(1) VT_BL(2) VT_BL(3) VT_BL(4) VT_BL(5) VT_BL(6) VT_BL(7) VT_BL(8)
  }
#undef VT_BL
  return 1;
}

}  // namespace vt
