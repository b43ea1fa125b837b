// vt_device.cuh -- device-side building blocks shared by the voltool kernels (sm_100a).
//
// All translation units are compiled with --fmad=false: every float expression below rounds
// after each operation exactly as the reference's scalar C++ does (built without contraction in
// the oracle).  Fused multiply-adds are written explicitly (__fmaf_rn / fma) where wanted.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vt_host.h"

namespace vt {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- context ------------------------------------------------------------------------------
struct Ctx {
  int device = -1;
  cudaStream_t stream = nullptr;
  int sm_count = kSMs;
  long launches = 0;
  double *d_scratch = nullptr;    // reduction partials
  size_t scratch_bytes = 0;
  unsigned int *d_ticket = nullptr;
  double *h_pinned = nullptr;     // small pinned readback area
  size_t pinned_bytes = 0;
};
Ctx &ctx();
bool ctx_ready();
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
cudaError_t last_cuda_error();   // the error of the most recent cuda_fail (to tell an allocation failure from a fault)

#define VT_CUDA(expr)                                                       \
  do {                                                                      \
    cudaError_t _e = (expr);                                                \
    if (_e != cudaSuccess) return ::vt::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)
#define VT_LAUNCHED()                                                       \
  do {                                                                      \
    ::vt::ctx().launches++;                                                 \
    cudaError_t _e = cudaGetLastError();                                    \
    if (_e != cudaSuccess) return ::vt::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)
// every entry point opens with this: the context check, then an NVTX range named after the function for as long as
// the call runs (the reference brackets its commands with PROFILE_PUSH_RANGE / PROFILE_POP_RANGE, TclMDFF.C:383,
// 551, 603; nested library calls nest their ranges).  nvtx3 is header-only and costs nothing without a tool attached.
#define VT_REQUIRE_CTX()                                                    \
  do {                                                                      \
    if (!::vt::ctx_ready()) {                                               \
      ::vt::set_error("voltool_gpu: vt_init() has not succeeded (no CUDA device, no CPU fallback)"); \
      return 1;                                                             \
    }                                                                       \
  } while (0);                                                              \
  ::vt::NvtxScope vt_nvtx_scope_(__func__)

struct NvtxScope {
  explicit NvtxScope(const char *name);
  ~NvtxScope();
};

// optional per-kernel CUDA-event timing on the library stream (bench.py's roofline leg)
enum ProfKey { PK_TRANSFORM = 0, PK_SETUP, PK_SPLAT, PK_POSE_CC, PK_POSE_FINISH, PK_CC, PK_STATS, PK_UNARY, PK_BINARY,
               PK_BLUR, PK_RESAMPLE, PK_COUNT };
void prof_begin(int key);
void prof_end(int key);

int dev_alloc(void **p, size_t bytes);
int dev_free(void *p);

// ---- resident map --------------------------------------------------------------------------
}  // namespace vt

struct vt_map {
  vt_grid grid;
  vt_cache cache;
  float *d = nullptr;  // device data, x fastest
  long n = 0;
};

namespace vt {

// ---- exact float helpers ---------------------------------------------------------------------
// C's (int) cast on x86-64 (cvttss2si) returns INT_MIN for NaN / out-of-range; CUDA's saturates
// and maps NaN to 0.  The reference's bounds tests depend on the former.
__device__ __forceinline__ int trunc_like_x86(float v) {
  return (v >= -2147483648.0f && v < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;
}

// Voltool.h:32-35 myisnan (bit test)
__device__ __forceinline__ bool is_nan_bits(float f) { return (__float_as_uint(f) << 1) > 0xff000000u; }
__device__ __forceinline__ float quiet_nan() { return __uint_as_float(0x7fc00000u); }

// x86 SSE scalar arithmetic hands back the first NaN operand (quieted), and 0xffc00000 when it has
// to make one up (inf-inf, 0*inf, 0/0); CUDA returns the canonical 0x7fffffff for both.  Where
// NaNs can legitimately flow through bit-exact ops this restores the reference's bits.  `a..d`
// are the operands in the order the reference expression consults them.
__device__ __forceinline__ float nan_like_x86(float r, float a, float b = 0.f, float c = 0.f, float d = 0.f) {
  if (!is_nan_bits(r)) return r;
  if (is_nan_bits(a)) return __uint_as_float(__float_as_uint(a) | 0x00400000u);
  if (is_nan_bits(b)) return __uint_as_float(__float_as_uint(b) | 0x00400000u);
  if (is_nan_bits(c)) return __uint_as_float(__float_as_uint(c) | 0x00400000u);
  if (is_nan_bits(d)) return __uint_as_float(__float_as_uint(d) | 0x00400000u);
  return __uint_as_float(0xffc00000u);
}

struct DevMap {  // lookup constants of one source map
  const float *__restrict__ d;
  float inv[16];
  int nx, ny, nz;
};

struct DevCoord {  // voxel index -> cartesian of an output grid
  double ox, oy, oz;
  float xd[3], yd[3], zd[3];
  int nx, ny, nz;
};

// VolumetricData::voxel_coord (VolumetricData.h:142-153) / Voltool.h:63-77:
// float(int*float) products promoted and summed in double on top of the double origin.
__device__ __forceinline__ void voxel_coord(const DevCoord &c, int gx, int gy, int gz, float &x, float &y, float &z) {
  const float fx = (float)gx, fy = (float)gy, fz = (float)gz;
  x = (float)(((c.ox + (double)(fx * c.xd[0])) + (double)(fy * c.yd[0])) + (double)(fz * c.zd[0]));
  y = (float)(((c.oy + (double)(fx * c.xd[1])) + (double)(fy * c.yd[1])) + (double)(fz * c.zd[1]));
  z = (float)(((c.oz + (double)(fx * c.xd[2])) + (double)(fy * c.yd[2])) + (double)(fz * c.zd[2]));
}

// Matrix4::multpoint3d with the precomputed inverse (VolumetricData.C:227)
__device__ __forceinline__ void cart_to_voxel(const float *__restrict__ m, float x, float y, float z, float &vx,
                                              float &vy, float &vz) {
  const float iw = 1.0f / (((x * m[3] + y * m[7]) + z * m[11]) + m[15]);
  const float u = iw * x, v = iw * y, w = iw * z;
  vx = ((u * m[0] + v * m[4]) + w * m[8]) + iw * m[12];
  vy = ((u * m[1] + v * m[5]) + w * m[9]) + iw * m[13];
  vz = ((u * m[2] + v * m[6]) + w * m[10]) + iw * m[14];
}

__device__ __forceinline__ int clampi(int v, int n) { return v > 0 ? (v < n ? v : n - 1) : 0; }  // voxel_value_safe :253-260

// VolumetricData::voxel_value_interpolate, VolumetricData.C:264-291
__device__ __forceinline__ float trilinear(const DevMap &m, float xv, float yv, float zv) {
  const int x = trunc_like_x86(xv), y = trunc_like_x86(yv), z = trunc_like_x86(zv);
  const float xf = xv - (float)x, yf = yv - (float)y, zf = zv - (float)z;
  const long sy = m.nx, sz = (long)m.nx * m.ny;
  const int x0 = clampi(x, m.nx), x1 = clampi(x + 1, m.nx);
  const long r00 = clampi(z, m.nz) * sz + clampi(y, m.ny) * sy;
  const long r10 = clampi(z, m.nz) * sz + clampi(y + 1, m.ny) * sy;
  const long r01 = clampi(z + 1, m.nz) * sz + clampi(y, m.ny) * sy;
  const long r11 = clampi(z + 1, m.nz) * sz + clampi(y + 1, m.ny) * sy;
  const float a0 = __ldg(m.d + r00 + x0), a1 = __ldg(m.d + r00 + x1);
  const float b0 = __ldg(m.d + r10 + x0), b1 = __ldg(m.d + r10 + x1);
  const float c0 = __ldg(m.d + r01 + x0), c1 = __ldg(m.d + r01 + x1);
  const float d0 = __ldg(m.d + r11 + x0), d1 = __ldg(m.d + r11 + x1);
  const float xa = a0 + xf * (a1 - a0);
  const float xb = b0 + xf * (b1 - b0);
  const float xc = c0 + xf * (c1 - c0);
  const float xe = d0 + xf * (d1 - d0);
  const float ya = xa + yf * (xb - xa);
  const float yb = xc + yf * (xe - xc);
  const float r = ya + zf * (yb - ya);
  if (is_nan_bits(r)) {  // rare: NaN voxels or inf arithmetic; restore the x86 payload
    const float v[8] = {a0, a1, b0, b1, c0, c1, d0, d1};
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (is_nan_bits(v[k])) return __uint_as_float(__float_as_uint(v[k]) | 0x00400000u);
    return __uint_as_float(0xffc00000u);
  }
  return r;
}

// voxel_value_interpolate_from_coord[_safe], VolumetricData.C:305-322 / 338-355
template <bool SAFE>
__device__ __forceinline__ float sample_interp(const DevMap &m, float x, float y, float z) {
  float vx, vy, vz;
  cart_to_voxel(m.inv, x, y, z, vx, vy, vz);
  const int gx = trunc_like_x86(vx), gy = trunc_like_x86(vy), gz = trunc_like_x86(vz);
  if (gx < 0 || gx >= m.nx || gy < 0 || gy >= m.ny || gz < 0 || gz >= m.nz) return SAFE ? 0.0f : quiet_nan();
  return trilinear(m, vx, vy, vz);
}

// voxel_value_from_coord[_safe], VolumetricData.C:295-301 / 327-333 (m.inv must be the shifted
// inverse).  The reference tests "index > 0", so voxel 0 reads as outside.
template <bool SAFE>
__device__ __forceinline__ float sample_nearest(const DevMap &m, float x, float y, float z) {
  float vx, vy, vz;
  cart_to_voxel(m.inv, x, y, z, vx, vy, vz);
  const int gx = trunc_like_x86(vx), gy = trunc_like_x86(vy), gz = trunc_like_x86(vz);
  long idx = -1;
  if (!(gx < 0 || gx >= m.nx || gy < 0 || gy >= m.ny || gz < 0 || gz >= m.nz))
    idx = (long)gz * m.nx * m.ny + (long)gy * m.nx + gx;
  if (idx > 0) return __ldg(m.d + idx);
  return SAFE ? 0.0f : quiet_nan();
}

// ---- packed fp32 pairs (sm_100: fma.rn.f32x2 -> FFMA2, both halves IEEE round-to-nearest like fmaf) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float x, float y) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 r, float &x, float &y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(r)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ---- reductions --------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-level sum of K doubles per thread; result valid in thread 0.  smem: K * 32 doubles.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double *smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; ++k) smem[k * 32 + warp] = v[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double t = lane < nw ? smem[k * 32 + lane] : 0.0;
      v[k] = warp_sum(t);
    }
  }
  __syncthreads();
}

// Deterministic grid-wide finish: every block stores its K partials; the last block to arrive
// (ticket) adds them in block order and writes out[0..K).  The ticket resets itself.
template <int K>
__device__ __forceinline__ void grid_finish(const double (&v)[K], double *partials, unsigned int *ticket,
                                            double *out, double *smem) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) partials[(size_t)blockIdx.x * K + k] = v[k];
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  // fixed order: thread t takes blocks t, t+blockDim, ...; then a fixed tree
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += partials[(size_t)b * K + k];
  block_sum<K>(acc, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = acc[k];
    *ticket = 0u;
  }
}

// ---- launch-side internal API (implemented across the .cu files) ------------------------------------
int k_cc(const vt_map *A, const vt_map *B, double thr, double sums[5], long *n);
int k_fill(float *d, long n, float v);
int k_minmax(const float *d, long n, float *mn, float *mx);
int k_sum(const float *d, long n, double *sum);
int k_sumsq_delta(const float *d, long n, float mean, double *sum);
int k_minmaxsum(const float *d, long n, float *mn, float *mx, double *sum);
enum UnaryOp { U_CLAMP, U_SCALE, U_ADD, U_RESCALE, U_SIGMA, U_BINMASK, U_MDFFPOT };
int k_unary(UnaryOp op, float *d, long n, float p0, float p1, float p2, double pd);
int k_binary(int op, int interp, int use_union, const vt_map *A, const vt_map *B, const vt_grid &og, float *out);
int k_pad_copy(const float *src, int ox, int oy, float *dst, const PadPlan &p);
int k_downsample(const float *src, int ox, int oy, int oz, float *dst);
int k_supersample(const float *src, int ox, int oy, int oz, float *dst);
int k_blur(const float *src, float *dst, float *tmp, int nx, int ny, int nz, const float *w, int R);
int k_vol_com(const vt_map *m, double out[4]);
int k_histogram(const float *d, long n, float mn, double binwidthinv, int nbins, unsigned long long *d_bins);
}  // namespace vt
struct vt_fit_ctx;
namespace vt {
// rotations x shifts with one splat per rotation (vt_fit.cu); cc_out[ir * ns + s] on the host
int fit_eval_rot_shifts(vt_fit_ctx *f, const float com[3], const float *rot_deg, long nrot, const float *shifts, int ns, float *cc_out);

}  // namespace vt
