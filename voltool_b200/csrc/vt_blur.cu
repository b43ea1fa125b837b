// vt_blur.cu -- separable 3-D Gaussian blur (VolumetricData::gaussian_blur, VolumetricData.C:977-997).
//
// The reference delegates the arithmetic to GaussianBlur<float> (Segmentation.h, not in the
// snapshot); the definition held to here is the one in oracle/ (x -> y -> z passes, taps
// -R..R in order, acc = fmaf(w, v, acc), clamp-to-edge, f32 between passes).
//
// Pass structure: the x pass stages each row segment (+halo) in shared memory; the y and z
// passes march a register window along the filtered axis so each input element is read from
// L2/HBM once per pass.
#include <climits>
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "vt_device.cuh"

namespace vt {

constexpr int kMaxTaps = 2 * 96 + 1;
__constant__ float c_taps[kMaxTaps];

// ---- x pass: one row per block-iteration, row staged through shared memory -------------------
__global__ void __launch_bounds__(256) blur_x_kernel(const float *__restrict__ in, float *__restrict__ out, int nx,
                                                     long rows, int R) {
  extern __shared__ float srow[];  // nx + 2R
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float *src = in + row * nx;
    for (int i = threadIdx.x; i < nx + 2 * R; i += blockDim.x) {
      int j = i - R;
      j = j < 0 ? 0 : (j > nx - 1 ? nx - 1 : j);
      srow[i] = __ldcs(src + j);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < nx; x += blockDim.x) {
      float acc = 0.0f;
      for (int k = 0; k <= 2 * R; ++k) acc = __fmaf_rn(c_taps[k], srow[x + k], acc);
      out[row * nx + x] = acc;
    }
    __syncthreads();
  }
}

// ---- y / z pass: lines along `stride`, threads across the contiguous x (and outer) dimension ----
// Each thread owns one line position and slides along the axis in segments of SEG outputs,
// reading SEG + 2R inputs per segment (clamped at the ends).
template <int SEG>
__global__ void __launch_bounds__(256) blur_axis_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                        int n_axis, long stride, long inner, long outer, int R) {
  // a "line" is identified by (o, i): base = o * (n_axis * stride) + i, i in [0, inner)
  const long lines = inner * outer;
  const int nseg = (n_axis + SEG - 1) / SEG;
  const long work = lines * nseg;
  const long step = (long)gridDim.x * blockDim.x;
  for (long w = (long)blockIdx.x * blockDim.x + threadIdx.x; w < work; w += step) {
    // consecutive threads -> consecutive i (coalesced); segments are the slowest index
    const long line = w % lines;
    const int seg = (int)(w / lines);
    const long o = line / inner, i = line - o * inner;
    const float *src = in + o * (n_axis * stride) + i;
    float *dst = out + o * (n_axis * stride) + i;
    const int a0 = seg * SEG;
    float acc[SEG];
#pragma unroll
    for (int s = 0; s < SEG; ++s) acc[s] = 0.0f;
    // input j = a0 - R + t contributes to output a0+s with tap index k = t - s
    for (int t = 0; t < SEG + 2 * R; ++t) {
      int j = a0 - R + t;
      j = j < 0 ? 0 : (j > n_axis - 1 ? n_axis - 1 : j);
      const float v = __ldg(src + (long)j * stride);
#pragma unroll
      for (int s = 0; s < SEG; ++s) {
        const int k = t - s;
        if (k >= 0 && k <= 2 * R) acc[s] = __fmaf_rn(c_taps[k], v, acc[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < SEG; ++s)
      if (a0 + s < n_axis) dst[(long)(a0 + s) * stride] = acc[s];
  }
}

// ---- compile-time radius: taps become immediate constant-bank operands, every loop is static,
// and each thread produces a run of outputs from one sliding window of inputs (each input is
// loaded once per thread instead of once per tap).  Tap order per output is unchanged.
template <int R, int XB>
__global__ void __launch_bounds__(256) blur_x_fixed_kernel(const float *__restrict__ in, float *__restrict__ out, int nx,
                                                           long rows, int rows_per_block) {
  extern __shared__ __align__(16) float srows[];  // rows_per_block x pitch
  const int segs = (nx + XB - 1) / XB;
  const int pitch = ((segs * XB + 2 * R + 3) & ~3) + 4;
  for (long r0 = (long)blockIdx.x * rows_per_block; r0 < rows; r0 += (long)gridDim.x * rows_per_block) {
    const int nr = (int)((rows - r0) < rows_per_block ? (rows - r0) : rows_per_block);
    for (int rr = 0; rr < nr; ++rr) {
      const float *src = in + (r0 + rr) * nx;
      float *d = srows + rr * pitch;
      for (int c = threadIdx.x; c < pitch; c += blockDim.x) {
        int j = c - R;
        j = j < 0 ? 0 : (j > nx - 1 ? nx - 1 : j);
        d[c] = __ldcs(src + j);
      }
    }
    __syncthreads();
    // consecutive threads take consecutive XB-wide segments of one row: with XB = 4 every float4
    // read below is a conflict-free 512-byte warp access
    for (int t = threadIdx.x; t < nr * segs; t += blockDim.x) {
      const int rr = t / segs, sg = t - rr * segs;
      const float *w = srows + rr * pitch + sg * XB;  // window: outputs sg*XB.. need w[0 .. XB+2R)
      float v[XB + 2 * R + 3];
#pragma unroll
      for (int q = 0; q < (XB + 2 * R + 3) / 4; ++q) {
        const float4 f = *reinterpret_cast<const float4 *>(w + 4 * q);
        v[4 * q + 0] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
      }
      float acc[XB];
#pragma unroll
      for (int s = 0; s < XB; ++s) {
        float a = 0.0f;
#pragma unroll
        for (int k = 0; k <= 2 * R; ++k) a = __fmaf_rn(c_taps[k], v[s + k], a);
        acc[s] = a;
      }
      float *o = out + (r0 + rr) * nx + sg * XB;
      if (XB == 4 && (nx & 3) == 0) {
        *reinterpret_cast<float4 *>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      } else {
#pragma unroll
        for (int s = 0; s < XB; ++s)
          if (sg * XB + s < nx) o[s] = acc[s];
      }
    }
    __syncthreads();
  }
}

// x pass without shared memory: a thread produces 4 consecutive outputs from a window of aligned
// float4 loads (consecutive lanes -> consecutive 16-byte pieces; the overlap between neighbouring
// windows is served by L1).  Segments whose window would cross a row end take clamped scalar loads.
template <int R>
__global__ void __launch_bounds__(256) blur_x_direct_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                            int nx, long rows) {
  constexpr int OFF = (4 - (R & 3)) & 3;            // (x0 - R) - first aligned x, for x0 % 4 == 0
  constexpr int NQ = (OFF + 4 + 2 * R + 3) / 4;      // float4 loads per window
  const int segs = (nx + 3) >> 2;
  const long work = rows * segs;
  const long step = (long)gridDim.x * blockDim.x;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += step) {
    const long row = t / segs;
    const int sg = (int)(t - row * segs), x0 = sg * 4;
    const float *src = in + row * nx;
    const int first = x0 - R - OFF;
    float v[4 * NQ];
    if (first >= 0 && first + 4 * NQ <= nx) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 f = __ldg(reinterpret_cast<const float4 *>(src + first) + q);
        v[4 * q + 0] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4 * NQ; ++i) {
        int j = first + i;
        j = j < 0 ? 0 : (j > nx - 1 ? nx - 1 : j);
        v[i] = __ldg(src + j);
      }
    }
    float acc[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      float a = 0.0f;
#pragma unroll
      for (int k = 0; k <= 2 * R; ++k) a = __fmaf_rn(c_taps[k], v[OFF + s + k], a);
      acc[s] = a;
    }
    float *o = out + row * nx + x0;
    if (x0 + 4 <= nx) *reinterpret_cast<float4 *>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else
#pragma unroll
      for (int s = 0; s < 4; ++s)
        if (x0 + s < nx) o[s] = acc[s];
  }
}

template <int R, int SEG>
__global__ void __launch_bounds__(256) blur_axis_fixed_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                              int n_axis, long stride, long inner, long outer) {
  const long lines = inner * outer;
  const int nseg = (n_axis + SEG - 1) / SEG;
  const long work = lines * nseg;
  const long step = (long)gridDim.x * blockDim.x;
  for (long w = (long)blockIdx.x * blockDim.x + threadIdx.x; w < work; w += step) {
    const long line = w % lines;
    const int seg = (int)(w / lines);
    const long o = line / inner, i = line - o * inner;
    const float *src = in + o * (n_axis * stride) + i;
    float *dst = out + o * (n_axis * stride) + i;
    const int a0 = seg * SEG;
    float acc[SEG];
#pragma unroll
    for (int s = 0; s < SEG; ++s) acc[s] = 0.0f;
#pragma unroll
    for (int t = 0; t < SEG + 2 * R; ++t) {
      int j = a0 - R + t;
      j = j < 0 ? 0 : (j > n_axis - 1 ? n_axis - 1 : j);
      const float v = __ldg(src + (long)j * stride);
#pragma unroll
      for (int s = 0; s < SEG; ++s)
        if (t - s >= 0 && t - s <= 2 * R) acc[s] = __fmaf_rn(c_taps[t - s], v, acc[s]);
    }
#pragma unroll
    for (int s = 0; s < SEG; ++s)
      if (a0 + s < n_axis) dst[(long)(a0 + s) * stride] = acc[s];
  }
}

// x and y passes of one plane tile in a single kernel: a CTA x-blurs the kXT x (kYT + 2R) rows the tile's
// y pass needs into shared memory (aligned float4 windows straight from global memory, as
// blur_x_direct_kernel), then each thread y-blurs 8 outputs of one column from a register window.
// The intermediate never leaves the SM, so the blur costs two passes over HBM instead of three; the
// arithmetic per output is the reference's (VolumetricData.C:946-1000: taps in order, x then y).
constexpr int kXT = 32, kYT = 64;

template <int R>
__global__ void __launch_bounds__(256) blur_xy_kernel(const float *__restrict__ in, float *__restrict__ out, int nx,
                                                      int ny, int nz, int xt, int yt, int zchunk) {
  constexpr int HR = kYT + 2 * R;
  constexpr int OFF = (4 - (R & 3)) & 3;  // left slack so that windows start 16-byte aligned
  constexpr int NQ = (OFF + 4 + 2 * R + 3) / 4;
  constexpr int NSEG = HR * (kXT / 4);
  constexpr int ITER = (NSEG + 255) / 256;
  __shared__ __align__(16) float sxbuf[2][HR * kXT];
  const int tid = threadIdx.x;
  const long w = blockIdx.x;
  const int tx0 = (int)(w % xt) * kXT, ty0 = (int)((w / xt) % yt) * kYT, z0 = (int)(w / ((long)xt * yt)) * zchunk;
  const int z1 = min(nz, z0 + zchunk);
  const long nxy = (long)nx * ny;
  // z-invariant description of this thread's x segments
  int soff[ITER];    // voxel offset of the window start inside a plane (row * nx + xs); xs may be < 0
  int sxs[ITER];     // xs, or INT_MIN when the segment does not exist
  int sdst[ITER];    // shared-memory float index of the 4 outputs
  bool edge[ITER];   // window crosses a row end
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int sg = tid + 256 * it;
    const int row = sg / (kXT / 4), sx0 = (sg - row * (kXT / 4)) * 4;
    int y = ty0 - R + row;
    y = y < 0 ? 0 : (y > ny - 1 ? ny - 1 : y);
    const int xs = tx0 + sx0 - R - OFF;
    sxs[it] = sg < NSEG ? xs : INT_MIN;
    soff[it] = y * nx + xs;
    sdst[it] = row * kXT + sx0;
    edge[it] = xs < 0 || xs + 4 * NQ > nx;
  }
  // y phase: a thread owns two adjacent columns (one packed pair) and four rows
  const int xp = (tid & 15) * 2, yb = (tid >> 4) * 4;
  const int gx = tx0 + xp;
  const bool colok = gx < nx;                      // nx is a multiple of 4: the pair is in or out together
  const int rows_ok = min(4, ny - (ty0 + yb));    // outputs of this thread that are inside the map
  const long ooff = (long)(ty0 + yb) * nx + gx;
  int buf = 0;
  for (int z = z0; z < z1; ++z, buf ^= 1) {
    const float *plane = in + (long)z * nxy;
    float *sx = sxbuf[buf];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      if (sxs[it] == INT_MIN) continue;
      float v[4 * NQ];
      if (!edge[it]) {
        const float4 *p = reinterpret_cast<const float4 *>(plane + soff[it]);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const float4 f = __ldg(p + q);
          v[4 * q + 0] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        }
      } else {
        // xs and nx are multiples of 4, so an aligned quad is wholly inside the row or wholly
        // outside: outside quads are a splat of the row's first / last voxel (clamp-to-edge)
        const float *rowp = plane + (soff[it] - sxs[it]);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int j = sxs[it] + 4 * q;
          const int jc = j < 0 ? 0 : (j > nx - 4 ? nx - 4 : j);
          float4 f = __ldg(reinterpret_cast<const float4 *>(rowp + jc));
          if (j < 0) f = make_float4(f.x, f.x, f.x, f.x);
          if (j > nx - 4) f = make_float4(f.w, f.w, f.w, f.w);
          v[4 * q + 0] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        }
      }
      float acc[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        float a = 0.0f;
#pragma unroll
        for (int k = 0; k <= 2 * R; ++k) a = __fmaf_rn(c_taps[k], v[OFF + s + k], a);
        acc[s] = a;
      }
      *reinterpret_cast<float4 *>(sx + sdst[it]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    __syncthreads();   // one barrier per plane: the other buffer is only rewritten after the next one
    f32x2 v[4 + 2 * R];
#pragma unroll
    for (int t = 0; t < 4 + 2 * R; ++t) v[t] = *reinterpret_cast<const f32x2 *>(sx + (yb + t) * kXT + xp);
    float *dst = out + (long)z * nxy + ooff;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      f32x2 a = pack2(0.0f, 0.0f);
#pragma unroll
      for (int k = 0; k <= 2 * R; ++k) a = fma2(pack2(c_taps[k], c_taps[k]), v[s + k], a);   // FFMA2, taps in order
      float a0, a1;
      unpack2(a, a0, a1);
      if (colok && s < rows_ok) *reinterpret_cast<float2 *>(dst + s * nx) = make_float2(a0, a1);
    }
  }
}

// z pass, four consecutive x per thread: SEG output planes from a sliding window of SEG + 2R input
// quads (taps in reference order per output).  blockIdx.y selects the plane segment.
template <int R, int SEG>
__global__ void __launch_bounds__(256) blur_z4_kernel(const float *__restrict__ in, float *__restrict__ out, long quads,
                                                      int nz) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= quads) return;
  const int a0 = blockIdx.y * SEG;
  const float4 *src = reinterpret_cast<const float4 *>(in) + q;
  float4 *dst = reinterpret_cast<float4 *>(out) + q;
  float4 acc[SEG];
#pragma unroll
  for (int s = 0; s < SEG; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < SEG + 2 * R; ++t) {
    int j = a0 - R + t;
    j = j < 0 ? 0 : (j > nz - 1 ? nz - 1 : j);
    const float4 v = __ldg(src + (long)j * quads);
#pragma unroll
    for (int s = 0; s < SEG; ++s)
      if (t - s >= 0 && t - s <= 2 * R) {
        const float c = c_taps[t - s];
        acc[s].x = __fmaf_rn(c, v.x, acc[s].x); acc[s].y = __fmaf_rn(c, v.y, acc[s].y);
        acc[s].z = __fmaf_rn(c, v.z, acc[s].z); acc[s].w = __fmaf_rn(c, v.w, acc[s].w);
      }
  }
#pragma unroll
  for (int s = 0; s < SEG; ++s)
    if (a0 + s < nz) dst[(long)(a0 + s) * quads] = acc[s];
}

// ---- y and z passes in one kernel --------------------------------------------------------------------
// A thread owns two adjacent columns (one packed pair) and two adjacent rows, and marches a chunk of z.
// Per input plane it reads the 2 + 2R x-blurred rows of its y window (8-byte loads; the rows shared with
// neighbouring threads come from L1), forms its two y-blurred pairs (taps in order, FFMA2), and scatters
// them into 2R+1 pending z outputs per pair that live in registers: input plane zin is tap k of output
// zin + R - k, so every output receives its taps in the order k = 0 .. 2R, exactly the chain of the
// gather form (acc = fma(w[k], v[clamp(z + k - R)], acc) from 0).  The output whose last tap just
// arrived is stored.  Accumulator slots rotate with zin; the rotation is resolved at compile time by
// dispatching on zin mod (2R+1).  Edge planes are replicated by feeding the clamped plane's values for
// the virtual planes zin < 0 and zin >= nz.  No shared memory, no barrier; HBM traffic is the
// intermediate once plus the output once.
template <int R, int S>
__device__ __forceinline__ void blur_z_scatter(f32x2 (&acc)[2][2 * R + 1], const f32x2 (&v)[2], const f32x2 (&w2)[2 * R + 1],
                                               bool emit, float2 *__restrict__ o0, float2 *__restrict__ o1, bool ok0, bool ok1) {
  constexpr int NT = 2 * R + 1;
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    constexpr int dummy = 0; (void)dummy;
    const int slot = (S + R - k + 2 * NT) % NT;   // slot of output zin + R - k, with slot(z) = z mod NT and S = zin mod NT
#pragma unroll
    for (int c = 0; c < 2; ++c) acc[c][slot] = fma2(w2[k], v[c], k == 0 ? pack2(0.0f, 0.0f) : acc[c][slot]);
  }
  if (emit) {   // output zin - R: its last tap (k = 2R) was just added
    constexpr int slot = (S - R + 2 * NT) % NT;
    float a, b;
    if (ok0) { unpack2(acc[0][slot], a, b); __stcs(o0, make_float2(a, b)); }
    if (ok1) { unpack2(acc[1][slot], a, b); __stcs(o1, make_float2(a, b)); }
  }
}

// one real branch per step (a switch, i.e. a jump table): an if-chain gets if-converted into predicated
// copies, which issues two scatter bodies per step
template <int R>
__device__ __forceinline__ void blur_z_dispatch(int s, f32x2 (&acc)[2][2 * R + 1], const f32x2 (&v)[2],
                                                const f32x2 (&w2)[2 * R + 1], bool emit, float2 *__restrict__ o0,
                                                float2 *__restrict__ o1, bool ok0, bool ok1) {
  constexpr int NT = 2 * R + 1;
#define VT_ZCASE(S) case S: if constexpr (S < NT) blur_z_scatter<R, (S < NT ? S : 0)>(acc, v, w2, emit, o0, o1, ok0, ok1); break;
  switch (s) {
    VT_ZCASE(0) VT_ZCASE(1) VT_ZCASE(2) VT_ZCASE(3) VT_ZCASE(4) VT_ZCASE(5) VT_ZCASE(6) VT_ZCASE(7) VT_ZCASE(8)
    VT_ZCASE(9) VT_ZCASE(10) VT_ZCASE(11) VT_ZCASE(12) VT_ZCASE(13) VT_ZCASE(14) VT_ZCASE(15) VT_ZCASE(16)
    default: break;
  }
#undef VT_ZCASE
}

constexpr int kYZx = 16, kYZy = 16;   // a CTA covers 16 column pairs x 16 row pairs = 32 x 32 voxels per plane
constexpr int kYZStages = 4;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4b(void *smem_dst, const void *gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int R>
__global__ void __launch_bounds__(kYZx * kYZy, 2) blur_yz_kernel(const float *__restrict__ in, float *__restrict__ out, int nx,
                                                                 int ny, int nz, int xt, int yt, int zchunk) {
  constexpr int NT = 2 * R + 1;
  constexpr int ROWS = 2 * kYZy + 2 * R;       // staged rows per plane: the tile's rows + the y halo
  constexpr int PITCH = 2 * kYZx;              // floats per staged row
  constexpr int CHUNKS = ROWS * (PITCH / 4);   // 16-byte pieces per plane
  constexpr int CIT = (CHUNKS + kYZx * kYZy - 1) / (kYZx * kYZy);
  __shared__ __align__(16) float tile[kYZStages][ROWS * PITCH];
  const int tid = threadIdx.x;
  const long w = blockIdx.x;
  const int tx = (int)(w % xt), ty = (int)((w / xt) % yt), zc = (int)(w / ((long)xt * yt));
  const int lx = tid % kYZx, ly = tid / kYZx;
  const int x0 = tx * PITCH, yb = ty * 2 * kYZy;   // tile origin
  const int xp = (x0 >> 1) + lx;                   // column pair: voxels 2 xp, 2 xp + 1 (nx is a multiple of 4)
  const int y0 = yb + 2 * ly;                      // rows y0, y0 + 1
  const int zlo = zc * zchunk, zhi = min(nz, zlo + zchunk);
  const bool ok0 = 2 * xp < nx && y0 < ny, ok1 = ok0 && y0 + 1 < ny;
  const int nx2 = nx >> 1;
  const long nxy = (long)nx * ny, nxy2 = (long)nx2 * ny;
  // planes this chunk reads: clamp(zlo - R) .. clamp(zhi - 1 + R), staged in order, kYZStages - 1 ahead
  const int pf = max(zlo - R, 0), pl = min(zhi - 1 + R, nz - 1);
  const int nplanes = pl - pf + 1;
  // this thread's 16-byte pieces of a plane tile: source address in plane pf (rows clamped to the map), slot
  unsigned long long src[CIT];
  int dsts[CIT];
#pragma unroll
  for (int i = 0; i < CIT; ++i) {
    const int ch = tid + kYZx * kYZy * i;
    const int row = ch / (PITCH / 4), q = ch - row * (PITCH / 4);
    int y = yb - R + row;
    y = y < 0 ? 0 : (y > ny - 1 ? ny - 1 : y);
    const int x = x0 + 4 * q;
    dsts[i] = (ch < CHUNKS && x < nx) ? row * PITCH + 4 * q : -1;
    src[i] = (unsigned long long)(in + ((long)pf * nxy + (long)y * nx + min(x, nx - 4)));
  }
  const unsigned long long pstep = (unsigned long long)nxy * sizeof(float);
  int staged = 0;
  auto stage_plane = [&]() {
    if (staged < nplanes) {
      float *buf = tile[staged % kYZStages];
#pragma unroll
      for (int i = 0; i < CIT; ++i) {
        if (dsts[i] >= 0) cp_async16(buf + dsts[i], reinterpret_cast<const void *>(src[i]));
        src[i] += pstep;
      }
    }
    cp_async_commit_group();
    ++staged;
  };
  f32x2 acc[2][NT];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int k = 0; k < NT; ++k) acc[c][k] = pack2(0.0f, 0.0f);
  f32x2 w2[NT];
#pragma unroll
  for (int k = 0; k < NT; ++k) w2[k] = pack2(c_taps[k], c_taps[k]);
  f32x2 v[2] = {pack2(0.0f, 0.0f), pack2(0.0f, 0.0f)};
  float2 *o0 = reinterpret_cast<float2 *>(out) + ((long)zlo * nxy2 + (long)y0 * nx2 + xp);
#pragma unroll
  for (int k = 0; k < kYZStages - 1; ++k) stage_plane();
  int consumed = 0;
  int s = ((zlo - R) % NT + NT) % NT;
  const int wbase = (2 * ly) * PITCH + 2 * lx;   // first window row of this thread inside a staged tile
  for (int zin = zlo - R; zin < zhi + R; ++zin) {
    const int zp = zin < 0 ? 0 : (zin > nz - 1 ? nz - 1 : zin);
    if (zp - pf == consumed) {                    // a new plane: uniform over the CTA
      cp_async_wait_group<kYZStages - 2>();       // this thread's pieces of the plane have landed ...
      __syncthreads();                            // ... and everybody's; the stage refilled below has been read by all
      const float *tp = tile[consumed % kYZStages] + wbase;
      f32x2 win[2 + 2 * R];
#pragma unroll
      for (int t = 0; t < 2 + 2 * R; ++t) win[t] = *reinterpret_cast<const f32x2 *>(tp + t * PITCH);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        f32x2 a = pack2(0.0f, 0.0f);
#pragma unroll
        for (int k = 0; k < NT; ++k) a = fma2(w2[k], win[c + k], a);
        v[c] = a;
      }
      stage_plane();
      ++consumed;
    }
    const bool emit = zin - R >= zlo;             // the output completed by this step (always < zhi inside the loop)
    blur_z_dispatch<R>(s, acc, v, w2, emit, o0, o0 + nx2, ok0, ok1);
    if (emit) o0 += nxy2;
    s = s + 1 == NT ? 0 : s + 1;
  }
}

// ---- all three passes in one kernel -------------------------------------------------------------------
// blur_yz_kernel with the x pass in front of it inside the CTA: the RAW source rows of the tile (+ y halo,
// + x halo rounded out to 16-byte pieces, edges replicated at staging time) are staged with cp.async
// kXYZStages - 1 planes ahead; per plane the CTA first x-blurs the staged rows of the NEXT plane into a
// double-buffered tile (a thread makes 8 consecutive outputs of one row from a 6 x 16-byte window, as four
// packed pairs: taps in order, FFMA2; the odd-aligned input pairs cost moves), then y-blurs and z-scatters
// the CURRENT plane exactly as blur_yz_kernel does.  One barrier per plane; HBM traffic is the source
// once plus the output once.
constexpr int kXYZStages = 4;
constexpr int kRawW = 2 * kYZx + 16;   // staged columns x0 - 8 .. x0 + 2 kYZx + 7
constexpr int kRawP = kRawW + 4;       // row pitch: 52 floats, so that the 16-byte window loads of two rows x four
                                       // segments (a quarter warp) fall into 32 different banks

template <int R>
__global__ void __launch_bounds__(kYZx * kYZy, 2) blur_xyz_kernel(const float *__restrict__ in, float *__restrict__ out, int nx,
                                                                  int ny, int nz, int xt, int yt, int zchunk) {
  static_assert(R >= 1 && R <= 8, "the x halo is staged 8 columns wide");
  constexpr int NT = 2 * R + 1;
  constexpr int ROWS = 2 * kYZy + 2 * R;
  constexpr int PITCH = 2 * kYZx;
  constexpr int NTHR = kYZx * kYZy;
  constexpr int CHUNKS = ROWS * (kRawW / 4);
  constexpr int CIT = (CHUNKS + NTHR - 1) / NTHR;
  constexpr int XUNITS = ROWS * (PITCH / 8);          // x-pass work items: one row x 8 outputs
  static_assert(XUNITS <= NTHR, "one x-pass item per thread");
  constexpr int XSTART = 4 * ((8 - R) / 4);            // first staged column of the (16-byte aligned) window, relative to 8 s
  constexpr int XOFF = (8 - R) % 4;                    // window position of the first tap of output 0
  constexpr int XNQ = (XOFF + 8 + 2 * R + 3) / 4;      // 16-byte loads per window
  __shared__ __align__(16) float raw[kXYZStages][ROWS * kRawP];
  __shared__ __align__(16) float xtile[2][ROWS * PITCH];
  const int tid = threadIdx.x;
  const long w = blockIdx.x;
  const int tx = (int)(w % xt), ty = (int)((w / xt) % yt), zc = (int)(w / ((long)xt * yt));
  const int lx = tid % kYZx, ly = tid / kYZx;
  const int x0 = tx * PITCH, yb = ty * 2 * kYZy;
  const int xp = (x0 >> 1) + lx;
  const int y0 = yb + 2 * ly;
  const int zlo = zc * zchunk, zhi = min(nz, zlo + zchunk);
  const bool ok0 = 2 * xp < nx && y0 < ny, ok1 = ok0 && y0 + 1 < ny;
  const int nx2 = nx >> 1;
  const long nxy = (long)nx * ny, nxy2 = (long)nx2 * ny;
  const int pf = max(zlo - R, 0), pl = min(zhi - 1 + R, nz - 1);
  const int nplanes = pl - pf + 1;
  // staging: 16-byte pieces of the raw tile; a piece left of x = 0 or right of x = nx - 1 is a splat of the row's
  // first / last voxel (clamp-to-edge; nx is a multiple of 4, so a piece is wholly inside or wholly outside)
  unsigned long long src[CIT];
  int dsts[CIT];       // slot, or -1: none
  bool splat[CIT];     // replicate *src instead of copying 16 bytes
#pragma unroll
  for (int i = 0; i < CIT; ++i) {
    const int ch = tid + NTHR * i;
    const int row = ch / (kRawW / 4), q = ch - row * (kRawW / 4);
    int y = yb - R + row;
    y = y < 0 ? 0 : (y > ny - 1 ? ny - 1 : y);
    const int x = x0 - 8 + 4 * q;
    dsts[i] = ch < CHUNKS ? row * kRawP + 4 * q : -1;
    splat[i] = x < 0 || x > nx - 4;
    const int xs = x < 0 ? 0 : (x > nx - 4 ? nx - 1 : x);
    src[i] = (unsigned long long)(in + ((long)pf * nxy + (long)y * nx + xs));
  }
  const unsigned long long pstep = (unsigned long long)nxy * sizeof(float);
  int staged = 0;
  auto stage_plane = [&]() {
    if (staged < nplanes) {
      float *buf = raw[staged % kXYZStages];
#pragma unroll
      for (int i = 0; i < CIT; ++i) {
        if (dsts[i] >= 0) {
          if (!splat[i]) cp_async16(buf + dsts[i], reinterpret_cast<const void *>(src[i]));
          else {   // four asynchronous 4-byte copies of the edge voxel: no load latency at staging time
#pragma unroll
            for (int j = 0; j < 4; ++j) cp_async4b(buf + dsts[i] + j, reinterpret_cast<const void *>(src[i]));
          }
        }
        src[i] += pstep;
      }
    }
    cp_async_commit_group();
    ++staged;
  };
  // x pass of staged plane c into xtile[c & 1]
  const int xrow = tid / (PITCH / 8), xseg = tid - xrow * (PITCH / 8);
  auto x_pass = [&](int c) {
    if (tid >= XUNITS) return;
    const float *rp = raw[c % kXYZStages] + xrow * kRawP + 8 * xseg + XSTART;
    float v[4 * XNQ];
#pragma unroll
    for (int q = 0; q < XNQ; ++q) {
      const float4 f = *reinterpret_cast<const float4 *>(rp + 4 * q);
      v[4 * q + 0] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
    }
    f32x2 a[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      f32x2 t = pack2(0.0f, 0.0f);
#pragma unroll
      for (int k = 0; k < NT; ++k) {
        const float tap = c_taps[k];
        t = fma2(pack2(tap, tap), pack2(v[XOFF + 2 * p + k], v[XOFF + 2 * p + k + 1]), t);
      }
      a[p] = t;
    }
    float o[8];
#pragma unroll
    for (int p = 0; p < 4; ++p) unpack2(a[p], o[2 * p], o[2 * p + 1]);
    float *xo = xtile[c & 1] + xrow * PITCH + 8 * xseg;
    *reinterpret_cast<float4 *>(xo) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4 *>(xo + 4) = make_float4(o[4], o[5], o[6], o[7]);
  };
  f32x2 acc[2][NT];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int k = 0; k < NT; ++k) acc[c][k] = pack2(0.0f, 0.0f);
  f32x2 w2[NT];
#pragma unroll
  for (int k = 0; k < NT; ++k) w2[k] = pack2(c_taps[k], c_taps[k]);
  f32x2 v[2] = {pack2(0.0f, 0.0f), pack2(0.0f, 0.0f)};
  float2 *o0 = reinterpret_cast<float2 *>(out) + ((long)zlo * nxy2 + (long)y0 * nx2 + xp);
#pragma unroll
  for (int k = 0; k < kXYZStages - 1; ++k) stage_plane();
  cp_async_wait_group<kXYZStages - 2>();
  __syncthreads();
  x_pass(0);
  int consumed = 0;
  int s = ((zlo - R) % NT + NT) % NT;
  const int wbase = (2 * ly) * PITCH + 2 * lx;
  for (int zin = zlo - R; zin < zhi + R; ++zin) {
    const int zp = zin < 0 ? 0 : (zin > nz - 1 ? nz - 1 : zin);
    if (zp - pf == consumed) {                    // a new plane: uniform over the CTA
      cp_async_wait_group<kXYZStages - 3>();      // this thread's pieces of raw plane consumed+1 have landed ...
      __syncthreads();                            // ... and everybody's; xtile[consumed & 1] is complete; the buffers
                                                  // written below were last read before this barrier
      if (consumed + 1 < nplanes) x_pass(consumed + 1);
      const float *tp = xtile[consumed & 1] + wbase;
      f32x2 win[2 + 2 * R];
#pragma unroll
      for (int t = 0; t < 2 + 2 * R; ++t) win[t] = *reinterpret_cast<const f32x2 *>(tp + t * PITCH);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        f32x2 a = pack2(0.0f, 0.0f);
#pragma unroll
        for (int k = 0; k < NT; ++k) a = fma2(w2[k], win[c + k], a);
        v[c] = a;
      }
      stage_plane();                              // raw plane consumed + kXYZStages - 1 into the stage x_pass(consumed) read
      ++consumed;
    }
    const bool emit = zin - R >= zlo;
    blur_z_dispatch<R>(s, acc, v, w2, emit, o0, o0 + nx2, ok0, ok1);
    if (emit) o0 += nxy2;
    s = s + 1 == NT ? 0 : s + 1;
  }
}

template <int R>
static int blur_one_pass(const float *src, float *dst, int nx, int ny, int nz) {
  Ctx &c = ctx();
  if ((nx & 3) != 0 || nx < 4 || (long)nx * ny > 0x7fffffffL) return 1;
  const int xt = (nx / 2 + kYZx - 1) / kYZx, yt = ((ny + 1) / 2 + kYZy - 1) / kYZy;
  // z chunks: every chunk re-reads and re-blurs 2R halo planes, but the CTAs are long-running and only two fit
  // an SM, so a few more than fill the machine balance better (measured on 512^3: 1 / 2 / 4 / 8 chunks ->
  // 0.61 / 0.54 / 0.51 / 0.53 ms)
  const long tiles = (long)xt * yt;
  int zchunks = (int)((27L * c.sm_count / 4 + tiles - 1) / tiles);
  const char *zc_env = getenv("VOLTOOL_BLUR_ZCHUNKS");
  if (zc_env) zchunks = atoi(zc_env);
  if (zchunks < 1) zchunks = 1;
  int zchunk = (nz + zchunks - 1) / zchunks;
  if (zchunk < 32) zchunk = nz < 32 ? nz : 32;
  zchunks = (nz + zchunk - 1) / zchunk;
  const long work = (long)xt * yt * zchunks;
  if (work > 0x7fffffffL) return 1;
  static bool attr = false;
  if (!attr) {
    VT_CUDA(cudaFuncSetAttribute(blur_xyz_kernel<R>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr = true;
  }
  blur_xyz_kernel<R><<<(unsigned int)work, kYZx * kYZy, 0, c.stream>>>(src, dst, nx, ny, nz, xt, yt, zchunk);
  VT_LAUNCHED();
  return 0;
}

// x pass into tmp (float4 windows), then the fused y+z pass into dst
template <int R>
static int blur_x_then_yz(const float *src, float *dst, float *tmp, int nx, int ny, int nz) {
  Ctx &c = ctx();
  if ((nx & 3) != 0 || nx < 4 || (long)nx * ny > 0x7fffffffL) return 1;
  const long rows = (long)ny * nz;
  blur_x_direct_kernel<R><<<c.sm_count * 8, 256, 0, c.stream>>>(src, tmp, nx, rows);
  VT_LAUNCHED();
  const int xt = (nx / 2 + kYZx - 1) / kYZx, yt = ((ny + 1) / 2 + kYZy - 1) / kYZy;
  // z chunks: enough CTAs for two waves of 2 per SM, but chunks of at least 64 planes (each adds 2R halo planes)
  int zchunks = (int)((4L * c.sm_count + (long)xt * yt - 1) / ((long)xt * yt));
  if (zchunks < 1) zchunks = 1;
  int zchunk = (nz + zchunks - 1) / zchunks;
  if (zchunk < 64) zchunk = nz < 64 ? nz : 64;
  zchunks = (nz + zchunk - 1) / zchunk;
  const long work = (long)xt * yt * zchunks;
  if (work > 0x7fffffffL) return 1;
  blur_yz_kernel<R><<<(unsigned int)work, kYZx * kYZy, 0, c.stream>>>(tmp, dst, nx, ny, nz, xt, yt, zchunk);
  VT_LAUNCHED();
  return 0;
}

// x+y fused pass into tmp, then the z pass into dst
template <int R>
static int blur_two_pass(const float *src, float *dst, float *tmp, int nx, int ny, int nz) {
  Ctx &c = ctx();
  const int xt = (nx + kXT - 1) / kXT, yt = (ny + kYT - 1) / kYT;
  const int zchunk = 16;
  const long work = (long)xt * yt * ((nz + zchunk - 1) / zchunk);
  // unaligned rows (or planes beyond int offsets): three-pass path
  if (work > 0x7fffffffL || (nx & 3) != 0 || nx < 4 || (long)nx * ny > 0x7fffffffL) return 1;
  blur_xy_kernel<R><<<(unsigned int)work, 256, 0, c.stream>>>(src, tmp, nx, ny, nz, xt, yt, zchunk);
  VT_LAUNCHED();
  constexpr int SEG = 16;
  const long quads = (long)nx * ny / 4;
  const dim3 gz((unsigned int)((quads + 255) / 256), (unsigned int)((nz + SEG - 1) / SEG));
  blur_z4_kernel<R, SEG><<<gz, 256, 0, c.stream>>>(tmp, dst, quads, nz);
  VT_LAUNCHED();
  return 0;
}

template <int R>
static int blur_fixed(const float *src, float *dst, float *tmp, int nx, int ny, int nz) {
  Ctx &c = ctx();
  constexpr int XB = 4, SEG = 16;
  const long rows = (long)ny * nz;
  const int segs = (nx + XB - 1) / XB;
  const int pitch = ((segs * XB + 2 * R + 3) & ~3) + 4;
  int rpb = 1024 / (segs > 0 ? segs : 1);  // ~4 segments per thread per batch
  if (rpb < 1) rpb = 1;
  if (rpb > 16) rpb = 16;
  size_t smem = sizeof(float) * (size_t)rpb * pitch;
  while (smem > 96 * 1024 && rpb > 1) { rpb /= 2; smem = sizeof(float) * (size_t)rpb * pitch; }
  if (smem > 200 * 1024) return 1;  // absurdly wide rows: caller falls back
  if (smem > 48 * 1024)
    VT_CUDA(cudaFuncSetAttribute(blur_x_fixed_kernel<R, XB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long blocks = (rows + rpb - 1) / rpb;
  int grid = (int)(blocks < (long)c.sm_count * 8 ? blocks : (long)c.sm_count * 8);
  if ((nx & 3) == 0) {  // rows 16-byte aligned: direct float4 windows
    blur_x_direct_kernel<R><<<c.sm_count * 8, 256, 0, c.stream>>>(src, dst, nx, rows);
  } else {
    blur_x_fixed_kernel<R, XB><<<grid, 256, smem, c.stream>>>(src, dst, nx, rows, rpb);
  }
  VT_LAUNCHED();
  const int g2 = c.sm_count * 8;
  blur_axis_fixed_kernel<R, SEG><<<g2, 256, 0, c.stream>>>(dst, tmp, ny, (long)nx, (long)nx, (long)nz);
  VT_LAUNCHED();
  blur_axis_fixed_kernel<R, SEG><<<g2, 256, 0, c.stream>>>(tmp, dst, nz, (long)nx * ny, (long)nx * ny, 1L);
  VT_LAUNCHED();
  return 0;
}

int k_blur_tma(const float *src, float *dst, int nx, int ny, int nz, const float *w, int R);

int k_blur(const float *src, float *dst, float *tmp, int nx, int ny, int nz, const float *w, int R) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  if (2 * R + 1 > kMaxTaps) { set_error("vt_gaussian_blur: sigma too large (radius %d > 96)", R); return 1; }
  VT_CUDA(cudaMemcpyToSymbolAsync(c_taps, w, sizeof(float) * (2 * R + 1), 0, cudaMemcpyHostToDevice, c.stream));
  const long rows = (long)ny * nz;
  int grid = (int)(rows < (long)c.sm_count * 8 ? rows : (long)c.sm_count * 8);
  if (grid < 1) grid = 1;
  // x: src -> dst ; y: dst -> tmp ; z: tmp -> dst
  prof_begin(PK_BLUR);
  const char *bmode = getenv("VOLTOOL_BLUR_MODE");
  // one pass over HBM with tensor-map TMA staging (vt_blur_tma.cu): the faster form for short kernels (R <= 3, i.e.
  // sigma <= 1 voxel; 512^3: 0.34 vs 0.41 ms at sigma 1, 0.29 vs 0.39 ms at sigma 0.6), where its per-plane
  // handshakes are not yet outweighed by the taps; longer kernels keep the two-kernel form unless asked ("tma")
  // ("two": the two-kernel form whatever the radius, for comparisons)
  if ((bmode && !strcmp(bmode, "tma")) || (!bmode && R <= 3 && (long)nx * ny * nz >= (1L << 22))) {
    const int rc = k_blur_tma(src, dst, nx, ny, nz, w, R);
    if (rc == 0) { prof_end(PK_BLUR); return 0; }
    if (rc > 1) return rc;
  }   // "yz": x pass + fused y/z pass; default: fused x/y pass + z pass
  if (bmode && !strcmp(bmode, "xyz")) {   // all three passes in one kernel
    int rc = 1;
    switch (R) {
      case 1: rc = blur_one_pass<1>(src, dst, nx, ny, nz); break;
      case 2: rc = blur_one_pass<2>(src, dst, nx, ny, nz); break;
      case 3: rc = blur_one_pass<3>(src, dst, nx, ny, nz); break;
      case 4: rc = blur_one_pass<4>(src, dst, nx, ny, nz); break;
      case 5: rc = blur_one_pass<5>(src, dst, nx, ny, nz); break;
      case 6: rc = blur_one_pass<6>(src, dst, nx, ny, nz); break;
      default: break;
    }
    if (rc == 0) { prof_end(PK_BLUR); return 0; }
    if (rc > 1) return rc;
  }
  if (bmode && !strcmp(bmode, "yz")) {
    int rc = 1;
    switch (R) {
      case 1: rc = blur_x_then_yz<1>(src, dst, tmp, nx, ny, nz); break;
      case 2: rc = blur_x_then_yz<2>(src, dst, tmp, nx, ny, nz); break;
      case 3: rc = blur_x_then_yz<3>(src, dst, tmp, nx, ny, nz); break;
      case 4: rc = blur_x_then_yz<4>(src, dst, tmp, nx, ny, nz); break;
      case 5: rc = blur_x_then_yz<5>(src, dst, tmp, nx, ny, nz); break;
      case 6: rc = blur_x_then_yz<6>(src, dst, tmp, nx, ny, nz); break;
      default: break;
    }
    if (rc == 0) { prof_end(PK_BLUR); return 0; }
    if (rc > 1) return rc;
  }
  if (!getenv("VOLTOOL_BLUR_GENERIC") && !getenv("VOLTOOL_BLUR_SEPARATE")) {
    int rc = 1;
    switch (R) {
      case 1: rc = blur_two_pass<1>(src, dst, tmp, nx, ny, nz); break;
      case 2: rc = blur_two_pass<2>(src, dst, tmp, nx, ny, nz); break;
      case 3: rc = blur_two_pass<3>(src, dst, tmp, nx, ny, nz); break;
      case 4: rc = blur_two_pass<4>(src, dst, tmp, nx, ny, nz); break;
      case 5: rc = blur_two_pass<5>(src, dst, tmp, nx, ny, nz); break;
      case 6: rc = blur_two_pass<6>(src, dst, tmp, nx, ny, nz); break;
      case 7: rc = blur_two_pass<7>(src, dst, tmp, nx, ny, nz); break;
      case 8: rc = blur_two_pass<8>(src, dst, tmp, nx, ny, nz); break;
      default: break;
    }
    if (rc == 0) { prof_end(PK_BLUR); return 0; }
    if (rc > 1) return rc;
  }
  if (!getenv("VOLTOOL_BLUR_GENERIC")) {
    int rc = 1;
    switch (R) {
      case 1: rc = blur_fixed<1>(src, dst, tmp, nx, ny, nz); break;
      case 2: rc = blur_fixed<2>(src, dst, tmp, nx, ny, nz); break;
      case 3: rc = blur_fixed<3>(src, dst, tmp, nx, ny, nz); break;
      case 4: rc = blur_fixed<4>(src, dst, tmp, nx, ny, nz); break;
      case 5: rc = blur_fixed<5>(src, dst, tmp, nx, ny, nz); break;
      case 6: rc = blur_fixed<6>(src, dst, tmp, nx, ny, nz); break;
      case 7: rc = blur_fixed<7>(src, dst, tmp, nx, ny, nz); break;
      case 8: rc = blur_fixed<8>(src, dst, tmp, nx, ny, nz); break;
      case 9: rc = blur_fixed<9>(src, dst, tmp, nx, ny, nz); break;
      case 10: rc = blur_fixed<10>(src, dst, tmp, nx, ny, nz); break;
      case 12: rc = blur_fixed<12>(src, dst, tmp, nx, ny, nz); break;
      default: break;
    }
    if (rc == 0) { prof_end(PK_BLUR); return 0; }
    if (rc > 1) return rc;
  }
  blur_x_kernel<<<grid, 256, sizeof(float) * (nx + 2 * R), c.stream>>>(src, dst, nx, rows, R);
  VT_LAUNCHED();
  const int g2 = c.sm_count * 8;
  blur_axis_kernel<8><<<g2, 256, 0, c.stream>>>(dst, tmp, ny, (long)nx, (long)nx, (long)nz, R);
  VT_LAUNCHED();
  blur_axis_kernel<8><<<g2, 256, 0, c.stream>>>(tmp, dst, nz, (long)nx * ny, (long)nx * ny, 1L, R);
  prof_end(PK_BLUR);
  VT_LAUNCHED();
  return 0;
}

}  // namespace vt
