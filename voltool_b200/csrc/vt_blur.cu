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
